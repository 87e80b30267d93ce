"""Fixtures of the reference's callers (tests/golden/callers/, written by tests/golden/make_caller_golden.py from the
UNMODIFIED ppo.py / utils.py / league_training.py / metrics_logger.py) and helpers shared by the tests that use them."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np
import torch

from hash_policy import HashPolicy, flag_targets
from helpers import compiled
from marl_ctf_development_b200 import experiment_env_config

CALLERS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "callers")

with open(os.path.join(CALLERS, "cases.json")) as _f:
    CASES = json.load(_f)
SEED = CASES["seed"]
ROLLOUT_FIELDS = ("grid_states", "metadata_states", "actions", "use_action_mask", "logprobs", "rewards", "dones", "values",
                  "next_grid_state", "next_metadata_state", "next_done")


def env_config(exp, overrides):
    ec = experiment_env_config(exp)
    ec.update(overrides)
    return ec


def policies(exp, overrides, salts, device="cpu"):
    ce = compiled(exp, **overrides)
    n_obs, n_meta = ce.n_channels * ce.GRID_SIZE**2, ce.meta_size
    return [HashPolicy(n_obs, n_meta, s, flag_targets(ce)).to(device) for s in salts]


def load_rollout(name) -> dict:
    """The eleven values of PPOTrainer.get_single_rollout per env, stacked like train_ppo's buffers (ppo.py:298-305,
    362-376): the first eight [T*apt, E, ...], next_* [E, ...]."""
    z = np.load(os.path.join(CALLERS, name + ".npz"))
    shape = tuple(int(x) for x in z["grid_shape"])
    out = {k: z[k] for k in ROLLOUT_FIELDS[1:]}
    out["grid_states"] = np.unpackbits(z["grid_bits"])[: int(np.prod(shape))].reshape(shape).astype(np.float32)
    out["next_grid_state"] = z["next_grid_state"].astype(np.float32)
    return out


def load_duel(name) -> dict:
    z = np.load(os.path.join(CALLERS, name + ".npz"))
    return {k: z[k] for k in z.files}


def load_duel_json(name) -> bytes:
    with gzip.open(os.path.join(CALLERS, name + ".json.gz"), "rb") as f:
        return f.read()


def load_league(name) -> dict:
    with open(os.path.join(CALLERS, name + ".json")) as f:
        return json.load(f)


def matrix_to_json(m) -> dict:
    return {f"{a}|{b}": float(v) for (a, b), v in m.items()}


def jsonable(obj):
    if isinstance(obj, dict):
        return {str(k): jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [jsonable(v) for v in obj]
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    return obj


# ------------------------------------------------------------------------------------------------------------------
# the caller loops restated on the CPU oracle (batched over envs) — themselves checked against the fixtures in
# tests/test_callers_fixtures.py, so that the oracle-side loops used elsewhere are pinned to the reference's callers too
# ------------------------------------------------------------------------------------------------------------------
def oracle_rollout(ce, B, seed, env_id_base, agent, opponent, train_team1, T):
    from oracle.ctf_oracle import OracleBatch

    orc = OracleBatch(ce, B, seed=seed, env_id_base=env_id_base)
    orc.reset()  # the rollout starts with env.reset() (ppo.py:57): episode 1
    N = ce.N_AGENTS
    team = 0 if train_team1 else 1
    mine = [i for i in range(N) if ce.AGENT_TEAMS[i] == team]
    rev = [ce.cfg.reversed_action[a] for a in range(9)]
    flags = torch.tensor([float(ce.AGENT_TYPE_ACTION_MASK[ce.AGENT_TYPES[i]]) for i in range(N)])
    rec = {k: [] for k in ("actions", "rewards", "values", "grid_states", "metadata_states", "use_action_mask", "logprobs")}
    d = None
    for _ in range(T):
        obs, meta = orc.observe()
        acts = np.zeros((B, N), dtype=np.uint8)
        for i in range(N):
            pol = agent if ce.AGENT_TEAMS[i] == team else opponent
            a, lp, _, v = pol.get_action_and_value(torch.from_numpy(obs[:, i]), torch.from_numpy(meta[:, i]), flags[i].expand(B))
            if i in mine:
                rec["actions"].append(a.numpy().astype(np.float32))
                rec["values"].append(v.reshape(B).numpy().copy())
                rec["logprobs"].append(lp.numpy().copy())
                rec["grid_states"].append(obs[:, i].copy())
                rec["metadata_states"].append(meta[:, i].copy())
                rec["use_action_mask"].append(np.full(B, float(flags[i]), dtype=np.float32))
            a = a.numpy()
            acts[:, i] = [rev[x] for x in a] if ce.AGENT_TEAMS[i] == 1 else a
        r, d = orc.step(acts)
        for i in mine:
            rec["rewards"].append(r[:, i].copy())
    obs, meta = orc.observe()
    out = {k: np.stack(v) for k, v in rec.items()}
    out["dones"] = np.zeros_like(out["rewards"])             # never written by the reference (ppo.py:53)
    out["next_grid_state"] = obs[:, min(mine)][:, None]          # per env (1, C, G, G) / (1, M) like the reference (ppo.py:116-118)
    out["next_metadata_state"] = meta[:, min(mine)][:, None]
    out["next_done"] = d.astype(np.float32).reshape(B, 1)
    return out


def oracle_duel(ce, B, seed, env_id_base, agent, opponent, max_steps):
    """utils.duel (utils.py:500-573) for B envs on the oracle. Returns (results, counters [B,13,N], captures, steps)."""
    from oracle.ctf_oracle import OracleBatch

    orc = OracleBatch(ce, B, seed=seed, env_id_base=env_id_base)
    orc.reset()
    N = ce.N_AGENTS
    rev = [ce.cfg.reversed_action[a] for a in range(9)]
    flags = torch.tensor([float(ce.AGENT_TYPE_ACTION_MASK[ce.AGENT_TYPES[i]]) for i in range(N)])
    step_count, done = 0, False
    while not done:
        step_count += 1
        obs, meta = orc.observe()
        acts = np.zeros((B, N), dtype=np.uint8)
        for i in range(N):
            pol = agent if ce.AGENT_TEAMS[i] == 0 else opponent
            a = pol.get_action_and_value(torch.from_numpy(obs[:, i]), torch.from_numpy(meta[:, i]), flags[i].expand(B))[0].numpy()
            acts[:, i] = [rev[x] for x in a] if ce.AGENT_TEAMS[i] == 1 else a
        _, d = orc.step(acts)
        done = bool(d.all()) or step_count > max_steps
    st = orc.state()
    caps = st["captures"]
    return np.sign(caps[:, 0] - caps[:, 1]), st["stats"], caps, st["step"]
