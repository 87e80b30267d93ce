"""Caller parity on the GPU: the batched adapters against outputs of the reference's
UNMODIFIED callers — PPOTrainer.get_single_rollout (ppo.py:31-131), utils.duel (utils.py:500-573), utils.duel_json
(utils.py:728-814), LeagueTrainer.calculate_winrate_matrix_* / generate_metrics_* (league_training.py:368-455,
573-648) — committed as fixtures (tests/golden/callers/, generator tests/golden/make_caller_golden.py), and the
reference's own callers run unmodified on the CUDA backend's single-env view where oracle/_ref travelled along."""
import json

import numpy as np
import pytest
import torch

import caller_cases as cc
from helpers import bits
from oracle import ref_shim as rs

pytestmark = pytest.mark.gpu
needs_reference = pytest.mark.skipif(not rs.available(), reason="reference neither under /root/reference nor in oracle/_ref")


@pytest.mark.parametrize("case", cc.CASES["rollouts"], ids=lambda c: c[0])
def test_collect_rollout_equals_unmodified_get_single_rollout(case):
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.rollout import collect_rollout

    name, exp, overrides, T, env_id0, n_envs, train_team1 = case
    want = cc.load_rollout(name)
    env = GridworldCtfGPU(**cc.env_config(exp, overrides), num_envs=n_envs, device="cuda:0", seed=cc.SEED, env_id_base=env_id0,
                          reverse_team1_actions=True)
    agent, opponent = cc.policies(exp, overrides, (1, 2), "cuda")
    ro = collect_rollout(env, agent, opponent, train_team1=train_team1, num_env_steps=T)
    got = dict(zip(cc.ROLLOUT_FIELDS, (ro.grid_states, ro.metadata_states, ro.actions, ro.use_action_mask, ro.logprobs,
                                       ro.rewards, ro.dones, ro.values)))
    for k, v in got.items():
        assert tuple(v.shape) == want[k].shape, k
        assert np.array_equal(bits(v.cpu().numpy()), bits(want[k])), k
    # per-env bootstrap values: the reference returns (1, C, G, G) / (1, M) / (1,) per rollout (ppo.py:111-118)
    assert np.array_equal(ro.next_grid_state.cpu().numpy(), want["next_grid_state"][:, 0])
    assert np.array_equal(bits(ro.next_metadata_state.cpu().numpy()), bits(want["next_metadata_state"][:, 0]))
    assert np.array_equal(ro.next_done.cpu().numpy(), want["next_done"][:, 0])
    # what train_ppo keeps (ppo.py:362-376): the first eight as they are, next_* of the LAST rollout only
    arrays = ro.as_reference_arrays()
    assert len(arrays) == 11
    assert np.array_equal(arrays[8].cpu().numpy(), want["next_grid_state"][-1])
    assert np.array_equal(bits(arrays[9].cpu().numpy()), bits(want["next_metadata_state"][-1]))
    assert np.array_equal(arrays[10].cpu().numpy(), want["next_done"][-1])


@pytest.mark.parametrize("case", cc.CASES["duels"], ids=lambda c: c[0])
def test_batched_duel_equals_unmodified_utils_duel(case):
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.config import METRIC_NAMES
    from marl_ctf_development_b200.rollout import batched_duel

    name, exp, overrides, max_steps, env_id0, n = case
    want = cc.load_duel(name)
    agent, opponent = cc.policies(exp, overrides, (3, 4), "cuda")
    env = GridworldCtfGPU(**cc.env_config(exp, overrides), num_envs=n, device="cuda:0", seed=cc.SEED, env_id_base=env_id0,
                          reverse_team1_actions=True, stats="counters")
    res = batched_duel(env, agent, opponent, max_steps=max_steps, return_result=True)               # utils.py:562-569
    assert np.array_equal(res.cpu().numpy(), want["results"])
    assert np.array_equal(env.counters().cpu().numpy(), want["counters"])                            # env.metrics of every duel
    assert np.array_equal(env.flag_captures().cpu().numpy(), want["captures"])
    assert np.array_equal(env.step_counts().cpu().numpy(), want["steps"])
    env = GridworldCtfGPU(**cc.env_config(exp, overrides), num_envs=n, device="cuda:0", seed=cc.SEED, env_id_base=env_id0,
                          reverse_team1_actions=True, stats="counters")   # a fresh env: the duel's reset starts episode 1 again
    metrics = batched_duel(env, agent, opponent, max_steps=max_steps, return_result=False)           # utils.py:571, summed over envs
    for k, mname in enumerate(METRIC_NAMES):
        for i in range(env.N_AGENTS):
            assert metrics["agent_" + mname][i] == int(want["counters"][:, k, i].sum())


@pytest.mark.parametrize("case", cc.CASES["duel_jsons"], ids=lambda c: c[0])
def test_duel_json_is_byte_identical_to_unmodified_utils_duel_json(case, tmp_path):
    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.trace_export import duel_json

    name, exp, overrides, max_steps, env_id = case
    want = cc.load_duel_json(name)
    agent, opponent = cc.policies(exp, overrides, (5, 6), "cuda")
    # the recorded env sits in the middle of a batch: env index 2 has global id env_id
    env = GridworldCtfGPU(**cc.env_config(exp, overrides), num_envs=5, device="cuda:0", seed=cc.SEED, env_id_base=env_id - 2,
                          reverse_team1_actions=True)
    path = tmp_path / "trace.json"
    duel_json(env, agent, opponent, env_index=2, max_steps=max_steps, fname=str(path))
    got = path.read_bytes()
    assert json.loads(got) == json.loads(want)
    assert got == want


@pytest.mark.parametrize("case", cc.CASES["leagues"], ids=lambda c: c[0])
def test_league_matrices_and_metrics_equal_the_unmodified_league_trainer(case):
    from marl_ctf_development_b200 import league

    name, exp, overrides, D, env_id0 = case
    want = cc.load_league(name)
    ec = cc.env_config(exp, overrides)
    pols = cc.policies(exp, overrides, (10, 11, 12, 20, 21, 22), "cuda")
    t1, t2 = pols[:3], pols[3:]
    kw = dict(device="cuda:0", seed=cc.SEED, env_id_base=env_id0)
    assert cc.matrix_to_json(league.winrate_matrix_symmetric(ec, t1, D, **kw)) == want["winrate_symmetric"]
    assert cc.matrix_to_json(league.winrate_matrix_non_symmetric(ec, t1, t2, D, **kw)) == want["winrate_non_symmetric"]
    if not rs.available():
        pytest.skip("MetricsLogger needs the reference (oracle/_ref): win-rate matrices checked, metrics harvest not")
    ML = rs.caller_modules()["metrics_logger"].MetricsLogger                # the reference's own logger, unmodified
    from helpers import compiled

    ce = compiled(exp, **overrides)
    team_types = {0: [], 1: []}
    for i in range(ce.N_AGENTS):
        team_types[ce.AGENT_TEAMS[i]].append(ce.AGENT_TYPES[i])
    metlog = ML(1, 1, 1, team_types, list(range(ce.N_AGENTS)), True)
    n = league.generate_metrics_symmetric(metlog, ec, t1, 0, D, 0, 1, 1, 1, **kw)
    assert n == 9 * D
    assert cc.jsonable(metlog.metrics) == want["metrics_symmetric"]
    metlog = ML(1, 1, 1, team_types, list(range(ce.N_AGENTS)), False)
    league.generate_metrics_non_symmetric(metlog, ec, t1, t2, 0, 0, D, 0, 1, 1, 1, **kw)
    assert cc.jsonable(metlog.metrics) == want["metrics_non_symmetric"]


# ------------------------------------------------------------------------------------------------------------------
# the reference's callers themselves, unmodified, driving the CUDA backend through the single-env view
# ------------------------------------------------------------------------------------------------------------------
@needs_reference
def test_unmodified_utils_duel_runs_on_the_cuda_backend():
    from marl_ctf_development_b200 import GridworldCtf

    duel = rs.caller_modules()["utils"].duel
    name, exp, overrides, max_steps, env_id0, n = cc.CASES["duels"][1]
    want = cc.load_duel(name)
    agent, opponent = cc.policies(exp, overrides, (3, 4))
    for k in (0, 3):
        env = GridworldCtf(**cc.env_config(exp, overrides), seed=cc.SEED, env_id=env_id0 + k)
        a, b, result = duel(env, agent, opponent, ("p", "q"), return_result=True, device="cpu", max_steps=max_steps)
        assert (a, b, result) == ("p", "q", int(want["results"][k]))
        assert env.env_step_count == int(want["steps"][k])
        env = GridworldCtf(**cc.env_config(exp, overrides), seed=cc.SEED, env_id=env_id0 + k)
        _, _, metrics = duel(env, agent, opponent, ("p", "q"), return_result=False, device="cpu", max_steps=max_steps)
        assert [metrics["team_flag_captures"][t] for t in (0, 1)] == want["captures"][k].tolist()
        from marl_ctf_development_b200.config import METRIC_NAMES

        for m, mname in enumerate(METRIC_NAMES):
            for i in range(env.N_AGENTS):
                assert metrics["agent_" + mname][i] == want["counters"][k, m, i], (mname, i)


@needs_reference
def test_unmodified_get_single_rollout_runs_on_the_cuda_backend():
    import types

    from marl_ctf_development_b200 import GridworldCtf

    PPOTrainer = rs.caller_modules()["ppo"].PPOTrainer
    name, exp, overrides, T, env_id0, n_envs, train_team1 = cc.CASES["rollouts"][3]    # 0_the_split, team 1 trained
    want = cc.load_rollout(name)
    agent, opponent = cc.policies(exp, overrides, (1, 2))
    k = 2
    env = GridworldCtf(**cc.env_config(exp, overrides), seed=cc.SEED, env_id=env_id0 + k)
    dims = env.get_env_dims()
    tr = PPOTrainer(types.SimpleNamespace(device="cpu", num_steps=T), dims[0], dims[2])
    tr.reverse_grid, tr.team_to_train, tr.device = (False, 0, "cpu") if train_team1 else (True, 1, "cpu")
    tr.num_agents_per_team = env.N_AGENTS // 2
    tr.num_steps = T * tr.num_agents_per_team
    tr.max_rewards = -np.inf
    out = tr.get_single_rollout(env, agent, opponent)
    for name_k, v in zip(cc.ROLLOUT_FIELDS[:8], out[:8]):
        assert np.array_equal(bits(v.numpy()), bits(want[name_k][:, k])), name_k
    assert np.array_equal(out[8].numpy(), want["next_grid_state"][k])
    assert np.array_equal(bits(out[9].numpy()), bits(want["next_metadata_state"][k]))
    assert np.array_equal(out[10].numpy(), want["next_done"][k])


@needs_reference
def test_unmodified_utils_duel_json_runs_on_the_cuda_backend(tmp_path):
    """utils.duel_json (utils.py:728-814), unmodified, on the single-env view: the file it writes is byte-identical to the
    one it wrote on the reference env (it reads env.grid, agent_positions, has_flag and metrics every step)."""
    from marl_ctf_development_b200 import GridworldCtf

    duel_json = rs.caller_modules()["utils"].duel_json
    name, exp, overrides, max_steps, env_id = cc.CASES["duel_jsons"][0]
    want = cc.load_duel_json(name)
    agent, opponent = cc.policies(exp, overrides, (5, 6))
    env = GridworldCtf(**cc.env_config(exp, overrides), seed=cc.SEED, env_id=env_id)
    path = tmp_path / "trace.json"
    duel_json(env, agent, opponent, max_steps=max_steps, fname=str(path))
    assert path.read_bytes() == want


@needs_reference
@pytest.mark.parametrize("exp", ["8_arena", "0_the_split"])
def test_unmodified_league_trainer_builds_its_objects_on_the_cuda_backend(exp, monkeypatch):
    """LeagueTrainer._init_objects (league_training.py:59-146) with the one-line import change of INTEGRATION.md §1: the env,
    its dims, the symmetry verdict, the per-team agent types and the Agent networks come out as with the reference env."""
    import marl_ctf_development_b200 as ours

    lt = rs.caller_modules()["league_training"]
    import runpy, os

    cfg_cls = runpy.run_path(os.path.join(rs.REF_ROOT, exp + rs.REF_SUFFIX), run_name="ctf_test")["TrainingConfig"]
    args = cfg_cls()
    with rs._in_ref_dir():
        ref_trainer = lt.LeagueTrainer(args)                       # the reference env (gridworld_ctf.GridworldCtf)
    monkeypatch.setattr(lt, "GridworldCtf", ours.GridworldCtf)      # "from marl_ctf_development_b200 import GridworldCtf"
    gpu_trainer = lt.LeagueTrainer(cfg_cls())
    assert isinstance(gpu_trainer.env, ours.GridworldCtf)
    for attr in ("local_grid_dims", "local_metadata_dims", "n_channels", "symmetric_teams"):
        assert getattr(gpu_trainer, attr) == getattr(ref_trainer, attr), attr
    assert dict(gpu_trainer.agent_team_types) == dict(ref_trainer.agent_team_types)
    assert gpu_trainer.metlog.agent_indiv_idxs == ref_trainer.metlog.agent_indiv_idxs
    a, b = gpu_trainer.main_agents_t1[0], ref_trainer.main_agents_t1[0]
    assert [tuple(p.shape) for p in a.parameters()] == [tuple(p.shape) for p in b.parameters()]
    # and the first thing train_league does with it: one unmodified duel between two of its agents
    torch.manual_seed(0)
    i, j, result = rs.caller_modules()["utils"].duel(gpu_trainer.env, a, gpu_trainer.main_agents_t1[-1], (0, 1), return_result=True,
                                                      device="cpu", max_steps=12)
    assert (i, j) == (0, 1) and result in (-1, 0, 1) and gpu_trainer.env.env_step_count == 13


@needs_reference
def test_reference_ppo_update_runs_on_batched_gpu_rollouts():
    """collect_rollout replaces the num_envs Ray rollouts of train_ppo (ppo.py:326-396); everything after —
    the reference's own calculate_advantages (:133-172) and optimise (:174-242), unmodified, with reference Agent networks —
    runs on its arrays because they have the same [num_steps * apt, num_envs, ...] layout."""
    import types

    import torch.optim as optim

    from marl_ctf_development_b200 import GridworldCtfGPU
    from marl_ctf_development_b200.rollout import collect_rollout

    mods = rs.caller_modules()
    Agent, PPOTrainer = mods["agent_network"].Agent, mods["ppo"].PPOTrainer
    ec = cc.env_config("8_arena", {"GAME_STEPS": 32})
    B, T = 16, 32
    env = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=5, reverse_team1_actions=True)
    dims = env.get_env_dims()
    torch.manual_seed(1)
    agent = Agent(9, env.n_channels, env.GRID_SIZE, env.meta_size, "cuda").cuda()
    opponent = Agent(9, env.n_channels, env.GRID_SIZE, env.meta_size, "cuda").cuda()
    args = types.SimpleNamespace(device="cuda", num_steps=T, num_envs=B, num_minibatches=4, update_epochs=2, gae=True, gamma=0.99,
                                 gae_lambda=0.95, norm_adv=True, clip_coef=0.2, clip_vloss=True, ent_coef=0.01, vf_coef=0.5,
                                 max_grad_norm=0.5, target_kl=None, learning_rate=2.5e-4)
    tr = PPOTrainer(args, dims[0], dims[2])
    apt = env.N_AGENTS // 2                                      # what train_ppo sets up (ppo.py:274-296)
    tr.device, tr.num_agents_per_team, tr.num_steps = "cuda", apt, T * apt
    tr.batch_size = B * T * apt
    tr.minibatch_size = tr.batch_size // args.num_minibatches
    tr.optimizer = optim.Adam(agent.parameters(), lr=args.learning_rate, eps=1e-5)
    before = [p.detach().clone() for p in agent.parameters()]
    for update in range(2):
        ro = collect_rollout(env, agent, opponent, train_team1=True)          # one 32-step episode of all 16 envs
        (grid_states, metadata_states, actions, use_action_mask, logprobs, rewards, dones, values,
         next_grid_state, next_metadata_state, next_done) = ro.as_reference_arrays()
        assert tuple(grid_states.shape) == (T * apt, B) + dims[0] and float(next_done) == 1.0
        advantages, returns = tr.calculate_advantages(agent, next_grid_state, next_metadata_state, rewards, next_done, dones, values)
        assert tuple(advantages.shape) == (T * apt, B) and bool(torch.isfinite(advantages).all())
        v_loss, pg_loss, ent = tr.optimise(agent, grid_states.reshape((-1,) + dims[0]), metadata_states.reshape((-1,) + dims[2]),
                                           logprobs.reshape(-1), actions.reshape(-1), use_action_mask.reshape(-1),
                                           advantages.reshape(-1), returns.reshape(-1), values.reshape(-1))
        assert all(np.isfinite(x) for x in (v_loss, pg_loss, ent)) and ent > 0
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.parameters()))
