"""Generates the committed fixtures from the UNMODIFIED reference (run where /root/reference exists):

    python tests/golden/make_golden.py

Writes
  marl_ctf_development_b200/data/experiments.json   env_config of the nine experiment scripts and of alt_exp/*.py (inputs)
  tests/golden/trace_<experiment>_<policy>.npz      per-step outputs of the reference GridworldCtf with the
                                                    Philox site draws injected (oracle/ref_shim.py)
  tests/golden/json_reset_states.json               reset-state content of the reference's json/*.json traces
                                                    (utils.py:745-754) for the experiments that have a script

The reference tree does not travel to the GPU box; these files do.
"""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import traces  # noqa: E402
from marl_ctf_development_b200.config import compile_config  # noqa: E402
from marl_ctf_development_b200.scenario_io import dump_env_config  # noqa: E402
from oracle import ref_shim as rs  # noqa: E402

SEED = 20260718
SECOND_EPISODE_STEPS = 40

# (experiment, policies): the three BASELINE.json configs get every policy, the rest the capture-heavy one
PLAN = {
    "0_the_split": ("uniform", "seek", "builder"),
    "7_gridlocked": ("uniform", "seek", "builder"),
    "8_arena": ("uniform", "seek", "builder"),
    "1_fence": ("seek",),
    "2_jailbreak": ("seek",),
    "3_one_way_out": ("seek",),
    "4_keyhole": ("builder",),
    "5_skittles": ("seek",),
    "6_the_wall": ("builder",),
}


def record(name: str, kind: str, env_id: int) -> dict:
    ec = rs.experiment_env_config(name)
    ce = compile_config(**ec)
    env = rs.make_injected_env(ec, seed=SEED, env_id=env_id)
    pol = traces.make_policy(kind, ce)
    rng = np.random.default_rng(env_id)
    T1 = ec["GAME_STEPS"] + 2  # two steps past done: done stays True, the count keeps increasing
    rec = {k: [] for k in ("actions", "grid", "pos", "hp_q", "has_flag", "inventory", "captures", "rewards", "done", "episode")}
    obs_steps, obs, meta = [], [], []
    stats, visits = [], []

    def snap_obs(t):
        o, m = rs.observations(env)
        obs_steps.append(t)
        obs.append(o.astype(np.uint8))
        meta.append(m)

    t = 0
    for episode, steps in ((0, T1), (1, SECOND_EPISODE_STEPS)):
        if episode:
            env.reset()
        snap_obs(t)  # observation of the reset state (index = number of steps recorded so far)
        for _ in range(steps):
            s = rs.snapshot(env, ce.cfg.hp_scale)
            a = pol(rng, s["pos"], s["has_flag"])
            _, r, d = env.step(a.tolist())
            s = rs.snapshot(env, ce.cfg.hp_scale)
            rec["actions"].append(a)
            for k in ("grid", "pos", "hp_q", "has_flag", "inventory", "captures"):
                rec[k].append(s[k])
            rec["rewards"].append(np.array(r, dtype=np.float32))
            rec["done"].append(np.uint8(d))
            rec["episode"].append(np.int32(episode))
            t += 1
            if traces.snap_after_step(t, s["step"], ec["GAME_STEPS"]):
                snap_obs(t)
        stats.append(rs.agent_metrics(env))
        visits.append(np.stack([env.metrics["agent_visitation_maps"][i] for i in range(ce.N_AGENTS)]))
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(
        seed=np.int64(SEED),
        env_id=np.int64(env_id),
        obs_steps=np.array(obs_steps, dtype=np.int32),
        obs=np.stack(obs),
        meta=np.stack(meta),
        stats=np.stack(stats),
        visits=np.stack(visits),
        episode_lengths=np.array([T1, SECOND_EPISODE_STEPS], dtype=np.int32),
    )
    return out


def json_reset_states() -> dict:
    out = {}
    for path in sorted(glob.glob(os.path.join(rs.REF_ROOT, "json", "*.json"))):
        base = os.path.basename(path)[: -len(".json")]
        exp = base.rsplit("_", 1)[0]
        exp = {"8_arena_iii": "8_arena"}.get(exp, exp)
        if exp not in rs.EXPERIMENTS:
            continue  # 2_jailbreak_ii_* have no experiment script in the tree
        with open(path) as f:
            d = json.load(f)
        out[base] = {
            "experiment": exp,
            "grid_size": d["grid_size"],
            "flag_pos": {k: [v["z"], v["x"]] for k, v in d["flag_pos"].items()},
            "spawn_pos": {k: [v["z"], v["x"]] for k, v in d["spawn_pos"].items()},
            "agents": [[a["team"], a["type"], a["start_z"], a["start_x"]] for a in d["agent_config"]],
            "block_tiles": sorted([t["z"], t["x"]] for t in d["block_tiles"]),
            "destructible_tiles": sorted([t["z"], t["x"], t["type"]] for t in d["destructible_tiles"]),
        }
    return out


def main():
    assert rs.available(), "needs /root/reference"
    exps = {name: dump_env_config(rs.experiment_env_config(name)) for name in rs.EXPERIMENTS}
    # the alternative experiment scripts (alt_exp/*.py: arena, arena_ii, jailbreak_ii, ... maps) as inputs only
    exps.update({"alt_exp/" + name: dump_env_config(rs.alt_experiment_env_config(name)) for name in rs.alt_experiment_names()})
    data_dir = os.path.join(ROOT, "marl_ctf_development_b200", "data")
    os.makedirs(data_dir, exist_ok=True)
    with open(os.path.join(data_dir, "experiments.json"), "w") as f:
        json.dump(exps, f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "json_reset_states.json"), "w") as f:
        json.dump(json_reset_states(), f, sort_keys=True)
    env_id = 100
    for name, kinds in PLAN.items():
        for kind in kinds:
            env_id += 1
            tr = record(name, kind, env_id)
            np.savez_compressed(os.path.join(HERE, f"trace_{name}_{kind}.npz"), **tr)
            print(name, kind, "captures", tr["captures"][tr["episode_lengths"][0] - 1].tolist(), "stats", tr["stats"][0].sum(1).tolist())


if __name__ == "__main__":
    main()
