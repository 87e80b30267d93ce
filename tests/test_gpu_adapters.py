"""Drop-in adapters on the GPU: the single-env view with the reference's surface, and the batched
rollout / duel loops against the same loops on the CPU oracle at larger batch sizes.  The loops
themselves are pinned to the reference's UNMODIFIED callers by tests/test_gpu_callers.py (fixtures generated from
ppo.py / utils.py / league_training.py) and tests/test_callers_fixtures.py (the oracle-side loops vs those fixtures)."""
import numpy as np
import pytest
import torch

import traces
from hash_policy import HashPolicy
from helpers import bits, compiled, golden_traces
from marl_ctf_development_b200 import experiment_env_config
from oracle.ctf_oracle import OracleBatch, OracleEnv

pytestmark = pytest.mark.gpu


def _reference_style_rollout(ce, B, seed, agent, opponent, train_team1, T):
    """ppo.py:31-131 per env on the oracle, with get_reversed_action applied on the host."""
    orc = OracleBatch(ce, B, seed=seed)
    orc.reset()  # the rollout starts with env.reset() (ppo.py:57): episode 1
    N = ce.N_AGENTS
    team = 0 if train_team1 else 1
    mine = [i for i in range(N) if ce.AGENT_TEAMS[i] == team]
    rev = [ce.cfg.reversed_action[a] for a in range(9)]
    flags = torch.tensor([float(ce.AGENT_TYPE_ACTION_MASK[ce.AGENT_TYPES[i]]) for i in range(N)])
    rec = {"actions": [], "rewards": [], "values": [], "grid": [], "meta": []}
    for t in range(T):
        obs, meta = orc.observe()
        acts = np.zeros((B, N), dtype=np.uint8)
        for i in range(N):
            pol = agent if ce.AGENT_TEAMS[i] == team else opponent
            a, _, _, v = pol.get_action_and_value(torch.from_numpy(obs[:, i]), torch.from_numpy(meta[:, i]), flags[i].expand(B))
            if i in mine:
                rec["actions"].append(a.numpy().copy())
                rec["values"].append(v.reshape(B).numpy().copy())
                rec["grid"].append(obs[:, i].copy())
                rec["meta"].append(meta[:, i].copy())
            a = a.numpy()
            acts[:, i] = [rev[x] for x in a] if ce.AGENT_TEAMS[i] == 1 else a
        r, d = orc.step(acts)
        for i in mine:
            rec["rewards"].append(r[:, i].copy())
    obs, meta = orc.observe()
    return {k: np.stack(v) for k, v in rec.items()}, obs[:, min(mine)], meta[:, min(mine)], d


@pytest.mark.parametrize("train_team1", [True, False])
def test_batched_rollout_matches_reference_loop(train_team1):
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.rollout import collect_rollout

    exp, B, seed, T = "8_arena", 48, 9, 70
    ec = experiment_env_config(exp)
    env = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=seed, reverse_team1_actions=True)
    n_obs, n_meta = env.n_channels * env.GRID_SIZE**2, env.meta_size
    agent, opponent = HashPolicy(n_obs, n_meta, 1), HashPolicy(n_obs, n_meta, 2)
    ro = collect_rollout(env, agent.cuda(), opponent.cuda(), train_team1=train_team1, num_env_steps=T)
    want, next_g, next_m, done = _reference_style_rollout(compiled(exp), B, seed, agent.cpu(), opponent.cpu(), train_team1, T)
    assert np.array_equal(ro.actions.cpu().numpy(), want["actions"])
    assert np.array_equal(bits(ro.rewards.cpu().numpy()), bits(want["rewards"]))
    assert np.array_equal(ro.values.cpu().numpy(), want["values"])
    assert np.array_equal(ro.grid_states.cpu().numpy(), want["grid"])
    assert np.array_equal(bits(ro.metadata_states.cpu().numpy()), bits(want["meta"]))
    assert np.array_equal(ro.next_grid_state.cpu().numpy(), next_g)
    assert np.array_equal(bits(ro.next_metadata_state.cpu().numpy()), bits(next_m))
    assert float(ro.dones.abs().sum()) == 0.0  # never written by the reference (ppo.py:53)
    apt = env.N_AGENTS // 2
    assert tuple(ro.grid_states.shape) == (T * apt, B, env.n_channels, env.GRID_SIZE, env.GRID_SIZE)
    arrays = ro.as_reference_arrays()   # ppo.py:298-305 / 362-376 order and shapes
    assert len(arrays) == 11 and tuple(arrays[8].shape) == (1, env.n_channels, env.GRID_SIZE, env.GRID_SIZE)
    assert tuple(arrays[5].shape) == (T * apt, B) and tuple(arrays[10].shape) == (1,)


def test_rollout_with_packed_observation_storage():
    """Packed rollout storage (1 bit per element) unpacks to exactly the float32 rollout."""
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.rollout import collect_rollout

    ec = experiment_env_config("8_arena")
    B, T = 24, 40
    n_obs, n_meta = 14 * 15 * 15, 22
    agent, opponent = HashPolicy(n_obs, n_meta, 1).cuda(), HashPolicy(n_obs, n_meta, 2).cuda()
    env_a = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=9, reverse_team1_actions=True)
    env_b = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=9, reverse_team1_actions=True, packed_obs=True)
    dense = collect_rollout(env_a, agent, opponent, num_env_steps=T)
    packed = collect_rollout(env_b, agent, opponent, num_env_steps=T, obs_storage_dtype="packed")
    assert packed.packed and packed.grid_states.dtype == torch.int32
    assert packed.grid_states.numel() * 4 * 31 < dense.grid_states.numel() * 4  # > 31x smaller
    assert torch.equal(packed.unpack_grid_states(env_b), dense.grid_states)
    mb = torch.tensor([5, 77, 120, 3], device="cuda")
    assert torch.equal(packed.unpack_grid_states(env_b, mb, dtype=torch.bfloat16).float(), dense.grid_states[mb])
    assert torch.equal(packed.actions, dense.actions) and torch.equal(packed.rewards, dense.rewards)


def test_batched_duel_matches_reference_loop():
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.rollout import batched_duel

    exp, B, seed, max_steps = "0_the_split", 64, 4, 90
    ec = experiment_env_config(exp)
    env = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=seed, reverse_team1_actions=True, stats="counters")
    ce = compiled(exp)
    n_obs, n_meta = env.n_channels * env.GRID_SIZE**2, env.meta_size
    agent, opponent = HashPolicy(n_obs, n_meta, 3), HashPolicy(n_obs, n_meta, 4)
    result = batched_duel(env, agent.cuda(), opponent.cuda(), max_steps=max_steps)
    # utils.duel on the oracle: reset, loop until done or step_count > max_steps
    orc = OracleBatch(ce, B, seed=seed)
    orc.reset()
    N = ce.N_AGENTS
    rev = [ce.cfg.reversed_action[a] for a in range(9)]
    flags = torch.tensor([float(ce.AGENT_TYPE_ACTION_MASK[ce.AGENT_TYPES[i]]) for i in range(N)])
    agent, opponent = agent.cpu(), opponent.cpu()
    step_count, done = 0, False
    while not done:
        step_count += 1
        obs, meta = orc.observe()
        acts = np.zeros((B, N), dtype=np.uint8)
        for i in range(N):
            pol = agent if ce.AGENT_TEAMS[i] == 0 else opponent
            a = pol.get_action(torch.from_numpy(obs[:, i]), torch.from_numpy(meta[:, i]), flags[i].expand(B)).numpy()
            acts[:, i] = [rev[x] for x in a] if ce.AGENT_TEAMS[i] == 1 else a
        _, d = orc.step(acts)
        done = bool(d.all()) or step_count > max_steps
    caps = orc.state()["captures"]
    assert np.array_equal(result.cpu().numpy(), np.sign(caps[:, 0] - caps[:, 1]))
    assert np.array_equal(env.stats_sum(all_reduce=False).cpu().numpy(), orc.state()["stats"].sum(0))


def test_ctf_policy_in_the_rollout_respects_masks_and_shapes():
    """BASELINE config 5 shape check: agent_network-style policy fed from the GPU env's buffers."""
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.policy import CtfPolicy
    from marl_ctf_development_b200.rollout import collect_rollout

    ec = experiment_env_config("8_arena")
    B = 32
    env = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=1, reverse_team1_actions=True)
    torch.manual_seed(0)
    agent = CtfPolicy(9, env.n_channels, env.GRID_SIZE, env.meta_size).cuda()
    opponent = CtfPolicy(9, env.n_channels, env.GRID_SIZE, env.meta_size).cuda()
    ro = collect_rollout(env, agent, opponent, train_team1=True, num_env_steps=20, obs_storage_dtype=torch.uint8)
    apt = 4
    assert tuple(ro.actions.shape) == (20 * apt, B) and ro.grid_states.dtype == torch.uint8
    masked_rows = ro.use_action_mask == 1  # guardians and scouts: only actions 0..4 (agent_network.py:66-75)
    assert bool((ro.actions[masked_rows] <= 4).all()) and bool(masked_rows.any()) and bool((~masked_rows).any())
    assert bool(torch.isfinite(ro.logprobs).all()) and bool(torch.isfinite(ro.values).all())


# ---------------------------------------------------------------------------------------------------
# single-env view with the reference's method surface
# ---------------------------------------------------------------------------------------------------
def test_single_env_view_replays_reference_trace_like_an_unmodified_caller():
    from marl_ctf_development_b200 import GridworldCtf
    from marl_ctf_development_b200.config import METRIC_NAMES

    exp, kind, path = [t for t in golden_traces() if t[0] == "7_gridlocked" and t[1] == "seek"][0]
    tr = np.load(path)
    env = GridworldCtf(**experiment_env_config(exp), seed=int(tr["seed"]), env_id=int(tr["env_id"]))
    assert env.get_env_dims() == ((13, 13, 13), (12, 13, 13), (18,), (87,))
    oi = 1
    T = int(tr["episode_lengths"][0])
    for t in range(T):
        grid, rewards, done = env.step([int(a) for a in tr["actions"][t]])   # list in, (grid, list, bool) out
        assert isinstance(rewards, list) and isinstance(done, bool)
        assert np.array_equal(grid, tr["grid"][t])
        assert np.array_equal(bits(np.array(rewards, dtype=np.float32)), bits(tr["rewards"][t]))
        assert done == bool(tr["done"][t])
        if t % 50 == 0:
            assert env.agent_positions == {i: tuple(int(x) for x in tr["pos"][t][i]) for i in range(env.N_AGENTS)}
            assert np.array_equal(env.has_flag, tr["has_flag"][t])
            assert [env.agent_hp[i] * 4 for i in range(env.N_AGENTS)] == tr["hp_q"][t].tolist()
        if traces.snap_after_step(t + 1, env.env_step_count, env.GAME_STEPS):
            assert tr["obs_steps"][oi] == t + 1
            for i in range(env.N_AGENTS):
                o = env.standardise_state(i, reverse_grid=(env.AGENT_TEAMS[i] != 0))
                assert o.dtype == np.uint8 and o.shape == (1, 13, 13, 13)
                assert np.array_equal(o[0], tr["obs"][oi][i])
                m = env.get_env_metadata(i)
                assert m.dtype == np.float16 and np.array_equal(m[0].astype(np.float32), tr["meta"][oi][i])
            oi += 1
    metrics = env.metrics
    for k, name in enumerate(METRIC_NAMES):
        for i in range(env.N_AGENTS):
            assert metrics["agent_" + name][i] == tr["stats"][0][k, i]
    assert metrics["team_flag_captures"][0] == tr["captures"][T - 1][0]
    assert np.array_equal(np.stack([metrics["agent_visitation_maps"][i] for i in range(env.N_AGENTS)]), tr["visits"][0])
    with pytest.raises(KeyError):
        env.step([9] * env.N_AGENTS)
    env.reset()
    assert env.env_step_count == 0 and not env.done


def test_single_env_view_symmetry_assert_and_reversed_actions():
    from marl_ctf_development_b200 import GridworldCtf

    env = GridworldCtf(**experiment_env_config("8_arena"))  # MAP_SYMMETRY_CHECK=True runs in the ctor (:476-477)
    assert np.array_equal(env.standardise_state(0), env.standardise_state(1, reverse_grid=True))
    assert [env.get_reversed_action(a) for a in range(9)] == [1, 0, 3, 2, 4, 6, 5, 8, 7]
    orc = OracleEnv(compiled("8_arena"))
    for i in range(8):
        for rev in (False, True):
            assert np.array_equal(env.standardise_state(i, reverse_grid=rev), orc.standardise_state(i, reverse_grid=rev))


def test_json_trace_export_matches_reference_wire_format(tmp_path):
    """utils.duel_json (utils.py:728-814): same keys / coordinate convention; content equals the same duel on the oracle."""
    import json

    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.trace_export import duel_json

    exp, B, seed, idx, max_steps = "7_gridlocked", 8, 6, 3, 150
    env = GridworldCtfGPU(**experiment_env_config(exp), num_envs=B, device="cuda:0", seed=seed, reverse_team1_actions=True)
    ce = compiled(exp)
    n_obs, n_meta = env.n_channels * env.GRID_SIZE**2, env.meta_size
    agent, opponent = HashPolicy(n_obs, n_meta, 5), HashPolicy(n_obs, n_meta, 6)
    path = tmp_path / "trace.json"
    duel_json(env, agent.cuda(), opponent.cuda(), env_index=idx, max_steps=max_steps, fname=str(path))
    d = json.loads(path.read_text())
    assert list(d) == ["grid_size", "flag_pos", "spawn_pos", "agent_config", "block_tiles", "destructible_tiles", "movement", "tiles", "scores"]
    assert len(d["movement"]) == max_steps + 1 == len(d["tiles"]) == len(d["scores"])  # 151 steps, like json/*.json
    assert d["flag_pos"]["0"] == {"x": ce.FLAG_POSITIONS[0][1], "z": ce.FLAG_POSITIONS[0][0]}
    # the same duel on the oracle, env `idx` only
    orc = OracleBatch(ce, 1, seed=seed, env_id_base=idx)
    orc.reset()
    N = ce.N_AGENTS
    rev = [ce.cfg.reversed_action[a] for a in range(9)]
    flags = torch.tensor([float(ce.AGENT_TYPE_ACTION_MASK[ce.AGENT_TYPES[i]]) for i in range(N)])
    agent, opponent = agent.cpu(), opponent.cpu()
    g0 = orc.state()["grid"][0]
    assert d["block_tiles"] == [{"x": int(x), "z": int(z)} for z, x in zip(*np.where(g0 == 1))]
    pos = orc.state()["pos"][0].astype(int)
    for t in range(max_steps + 1):
        obs, meta = orc.observe()
        acts = np.zeros((1, N), dtype=np.uint8)
        for i in range(N):
            pol = agent if ce.AGENT_TEAMS[i] == 0 else opponent
            a = int(pol.get_action(torch.from_numpy(obs[:, i]), torch.from_numpy(meta[:, i]), flags[i].expand(1)))   # one sample: a Python int, like Agent.get_action
            acts[0, i] = rev[a] if ce.AGENT_TEAMS[i] == 1 else a
        orc.step(acts)
        st = orc.state()
        new = st["pos"][0].astype(int)
        want = [{"x": int(new[i, 1] - pos[i, 1]), "z": int(new[i, 0] - pos[i, 0]), "has_flag": int(st["has_flag"][0][i])} for i in range(N)]
        assert d["movement"][t] == want, t
        assert d["scores"][t] == [{"t0": int(st["captures"][0][0]), "t1": int(st["captures"][0][1])}]
        g = st["grid"][0]
        assert len(d["tiles"][t]) == int(((g == 2) | (g == 3)).sum())
        pos = new


def test_league_duel_batches_match_per_pair_duels():
    """N3: all (agent, opponent) pairs of a win-rate matrix in one env batch == one utils.duel per pair on the oracle."""
    from marl_ctf_development_b200.league import duel_pairs, winrate_matrix_non_symmetric

    exp, D, seed, max_steps = "0_the_split", 6, 8, 60
    ec = experiment_env_config(exp)
    ce = compiled(exp)
    n_obs, n_meta = ce.n_channels * ce.GRID_SIZE**2, ce.meta_size
    t1 = [HashPolicy(n_obs, n_meta, 10 + i) for i in range(2)]
    t2 = [HashPolicy(n_obs, n_meta, 20 + i) for i in range(2)]
    pairs = [(a, o) for a in t1 for o in t2]
    results, metrics = duel_pairs(ec, [(a.cuda(), o.cuda()) for a, o in pairs], D, max_steps=max_steps, device="cuda:0", seed=seed, collect_metrics=True)
    N = ce.N_AGENTS
    rev = [ce.cfg.reversed_action[a] for a in range(9)]
    flags = torch.tensor([float(ce.AGENT_TYPE_ACTION_MASK[ce.AGENT_TYPES[i]]) for i in range(N)])
    for p, (agent, opponent) in enumerate(pairs):
        agent, opponent = agent.cpu(), opponent.cpu()
        orc = OracleBatch(ce, D, seed=seed, env_id_base=p * D)
        orc.reset()
        for step in range(max_steps + 1):
            obs, meta = orc.observe()
            acts = np.zeros((D, N), dtype=np.uint8)
            for i in range(N):
                pol = agent if ce.AGENT_TEAMS[i] == 0 else opponent
                a = pol.get_action(torch.from_numpy(obs[:, i]), torch.from_numpy(meta[:, i]), flags[i].expand(D)).numpy()
                acts[:, i] = [rev[x] for x in a] if ce.AGENT_TEAMS[i] == 1 else a
            orc.step(acts)
        so = orc.state()
        assert np.array_equal(results[p].cpu().numpy(), np.sign(so["captures"][:, 0] - so["captures"][:, 1]))
        want = so["stats"].sum(0)
        for k, name in enumerate(["tag_count", "respawn_tag_count", "flag_pickups", "flag_captures"]):
            for i in range(N):
                assert metrics[p]["agent_" + name][i] == want[k, i]
    wm = winrate_matrix_non_symmetric(ec, [a.cuda() for a in t1], [o.cuda() for o in t2], D, max_steps=max_steps, device="cuda:0", seed=seed)
    assert set(wm) == {(f"0_{i}", f"1_{j}") for i in range(2) for j in range(2)} | {(f"1_{j}", f"0_{i}") for i in range(2) for j in range(2)}
    r = results.cpu().numpy()
    assert abs(wm[("0_1", "1_0")] - (r[2] == 1).mean()) < 1e-12
    assert abs(wm[("1_0", "0_1")] - (r[2] == -1).mean()) < 1e-12


def test_bench_contract_keys():
    """bench.py prints one JSON line with the keys the driver reads (small sizes here; the real run uses the defaults)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    proc = subprocess.run(
        [sys.executable, os.path.join(root, "bench.py"), "--steps", "6", "--warmup", "3", "--envs", "2048", "--no-cpu-baseline"],
        capture_output=True, text=True, timeout=600, cwd=root,
    )
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["metric"] == "agent_steps_per_sec" and d["steps"] == 6 and d["gpu_launches"] >= 6 and d["value"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] == 2048 * 8
    assert "workload" in d["config"] and "model" not in d["config"]
    # BASELINE configs[1], [2] and [4] ride along in the same line (side_configs, config5_rollout)
    assert len(d["side_configs"]) == 2 and all(s["ms_per_step"] > 0 for s in d["side_configs"])
    assert d["config5_rollout"]["value"] > 0 and d["config5_rollout"]["env_only_value"] > d["config5_rollout"]["value"]
    assert d["episode_stats_checksum"] != 0 and d["clocks"]["samples"] >= 2


def test_integration_md_ctypes_stub_runs_as_written():
    """The reference-side binding shown in INTEGRATION.md §3 is executed verbatim (only the library path is made absolute)."""
    import os
    import re

    from marl_ctf_development_b200 import _native

    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if b.startswith("import ctypes as C, torch")]
    assert len(stub) == 1
    code = stub[0].replace('C.CDLL("libctf_b200.so")', f'C.CDLL({_native.LIB_PATH!r})')
    scope = {"env_config": experiment_env_config("8_arena")}
    exec(compile(code, "INTEGRATION.md", "exec"), scope)
    torch.cuda.synchronize()
    obs, rew, envs = scope["obs"], scope["rew"], scope["envs"]
    assert int(envs[:, 0].min()) == 1 and int(envs[:, 0].max()) == 1          # every env advanced one step
    assert bool((obs[:, :, 0].sum((2, 3)) == 1).all()) and bool(torch.isfinite(rew).all())
    # and it computed the same thing as the packaged binding
    from marl_ctf_development_b200 import GridworldCtfGPU

    env = GridworldCtfGPU(**experiment_env_config("8_arena"), num_envs=65536, device="cuda:0", seed=42)
    env.step(scope["actions"])
    assert torch.equal(env.obs, obs) and torch.equal(env.rewards, rew)
