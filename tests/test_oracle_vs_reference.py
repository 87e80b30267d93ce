"""Pins the CPU oracle against the imported, unmodified reference with injected draws.

Runs only where /root/reference exists (this container); the GPU box relies on the
committed traces in tests/golden/ instead.
"""
import numpy as np
import pytest

import traces
from helpers import STATE_KEYS, assert_state_equal, bits
from marl_ctf_development_b200.config import compile_config, env_dims
from oracle import ref_shim as rs
from oracle.ctf_oracle import OracleEnv
from kwarg_cases import CASE_IDS, KWARG_CASES, build_env_config

pytestmark = [
    pytest.mark.reference,
    pytest.mark.skipif(not rs.available(), reason="reference tree not present"),
    pytest.mark.filterwarnings("ignore"),
]

CASES = [
    ("0_the_split", "seek", 260),
    ("1_fence", "builder", 200),
    ("2_jailbreak", "seek", 260),
    ("3_one_way_out", "builder", 200),
    ("4_keyhole", "seek", 200),
    ("5_skittles", "builder", 200),
    ("6_the_wall", "seek", 200),
    ("7_gridlocked", "builder", 503),
    ("8_arena", "seek", 503),
    ("8_arena", "uniform", 200),
]


@pytest.mark.parametrize("exp,kind,steps", CASES, ids=[f"{e}-{k}" for e, k, _ in CASES])
def test_oracle_equals_reference_step_by_step(exp, kind, steps):
    ec = rs.experiment_env_config(exp)
    ce = compile_config(**ec)
    seed, env_id = 99, 1234
    ref = rs.make_injected_env(ec, seed=seed, env_id=env_id)
    orc = OracleEnv(ce, seed=seed, env_id=env_id)
    assert env_dims(ce) == ref.get_env_dims()
    assert [int(t) for t in ce.TILES_USED] == [int(t) for t in ref.TILES_USED]
    pol = traces.make_policy(kind, ce)
    rng = np.random.default_rng(5)
    for episode in range(2):
        if episode:
            ref.reset()
            orc.reset()
        for t in range(steps if episode == 0 else 30):
            s = rs.snapshot(ref, ce.cfg.hp_scale)
            a = pol(rng, s["pos"], s["has_flag"])
            _, rr, rd = ref.step(a.tolist())
            orr, od = orc.step(a)
            assert_state_equal(orc.state(), rs.snapshot(ref, ce.cfg.hp_scale), f"{exp}/{kind} ep{episode} t={t}")
            assert rd == od
            assert np.array_equal(bits(np.array(rr, dtype=np.float32)), bits(orr))
            if t % 5 == 0:
                ro, rm = rs.observations(ref)
                oo, om = orc.observe()
                assert np.array_equal(ro, oo)
                assert np.array_equal(bits(rm), bits(om))
        st = orc.state()
        assert np.array_equal(rs.agent_metrics(ref), st["stats"])
        vm = np.stack([ref.metrics["agent_visitation_maps"][i] for i in range(ce.N_AGENTS)])
        assert np.array_equal(vm, st["visits"])


def test_arbitrary_reverse_flags_and_symmetry_assert():
    """standardise_state(i, reverse_grid) for both flag values; the ctor's symmetry assert (gridworld_ctf.py:476-477)."""
    ec = rs.experiment_env_config("8_arena")
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=1, env_id=1)
    orc = OracleEnv(ce, seed=1, env_id=1)
    assert np.array_equal(orc.standardise_state(0), orc.standardise_state(1, reverse_grid=True))
    rng = np.random.default_rng(0)
    for t in range(60):
        a = traces.uniform_actions(rng, ce.N_AGENTS)
        ref.step(a.tolist())
        orc.step(a)
    for i in range(ce.N_AGENTS):
        for rev in (False, True):
            assert np.array_equal(ref.standardise_state(i, reverse_grid=rev), orc.standardise_state(i, reverse_grid=rev))


@pytest.mark.parametrize("exp,flip", [("0_the_split", 2), ("1_fence", 0), ("8_arena", None), ("7_gridlocked", 1)])
def test_all_flip_axes_are_covered(exp, flip):
    """FLIP_AXIS 2, 0 and None appear in shipped scenarios; 1 is forced onto a map to cover np.flip(axis=1)."""
    ec = rs.experiment_env_config(exp)
    ec["SCENARIO"] = dict(ec["SCENARIO"], FLIP_AXIS=flip)
    ec["MAP_SYMMETRY_CHECK"] = False
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=3, env_id=3)
    orc = OracleEnv(ce, seed=3, env_id=3)
    rng = np.random.default_rng(1)
    for t in range(25):
        a = traces.uniform_actions(rng, ce.N_AGENTS)
        ref.step(a.tolist())
        orc.step(a)
    for i in range(ce.N_AGENTS):
        assert np.array_equal(ref.standardise_state(i, reverse_grid=True), orc.standardise_state(i, reverse_grid=True))
    for a in range(9):
        assert ref.get_reversed_action(a) == ce.cfg.reversed_action[a]


@pytest.mark.parametrize("name,exp,overrides,kind", KWARG_CASES, ids=CASE_IDS)
def test_constructor_keyword_variations(name, exp, overrides, kind):
    """Every ctor keyword the step path reads (gridworld_ctf.py:19-52), odd team layouts, other HP grids."""
    ec = build_env_config(rs.experiment_env_config(exp), overrides)
    ce = compile_config(**ec)
    seed, env_id = 4242, 17
    ref = rs.make_injected_env(ec, seed=seed, env_id=env_id)
    orc = OracleEnv(ce, seed=seed, env_id=env_id)
    assert env_dims(ce) == ref.get_env_dims()
    assert [int(t) for t in ce.TILES_USED] == [int(t) for t in ref.TILES_USED]
    assert ce.OPPONENTS == ref.OPPONENTS
    pol = traces.make_policy(kind, ce)
    rng = np.random.default_rng(11)
    steps = min(ec["GAME_STEPS"] + 2, 300)
    for t in range(steps):
        s = rs.snapshot(ref, ce.cfg.hp_scale)
        a = pol(rng, s["pos"], s["has_flag"])
        _, rr, rd = ref.step(a.tolist())
        orr, od = orc.step(a)
        assert_state_equal(orc.state(), rs.snapshot(ref, ce.cfg.hp_scale), f"{name} t={t}", float_hp=bool(ce.cfg.hp_float))
        assert rd == od
        assert np.array_equal(bits(np.array(rr, dtype=np.float32)), bits(orr)), (name, t, rr, orr)
        if t % 7 == 0 or t == steps - 1:
            ro, rm = rs.observations(ref)
            oo, om = orc.observe()
            assert np.array_equal(ro, oo), (name, t)
            assert np.array_equal(bits(rm), bits(om)), (name, t, rm, om)
    st = orc.state()
    assert np.array_equal(rs.agent_metrics(ref), st["stats"])
    vm = np.stack([ref.metrics["agent_visitation_maps"][i] for i in range(ce.N_AGENTS)])
    assert np.array_equal(vm, st["visits"])


@pytest.mark.parametrize("seed", list(range(24)) + list(range(100, 108)))
def test_random_scenarios_and_configs(seed):
    """Random maps (grid 6..16, any FLIP_AXIS), team layouts, HP tables: reference == oracle step by step."""
    from random_scenarios import random_env_config

    ec = random_env_config(seed)
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=77, env_id=seed)
    orc = OracleEnv(ce, seed=77, env_id=seed)
    assert env_dims(ce) == ref.get_env_dims()
    assert [int(t) for t in ce.TILES_USED] == [int(t) for t in ref.TILES_USED], "TILES_USED order (set iteration) differs"
    pol = traces.make_policy("builder", ce)
    rng = np.random.default_rng(seed)
    for t in range(ec["GAME_STEPS"] + 2):
        s = rs.snapshot(ref, ce.cfg.hp_scale)
        a = pol(rng, s["pos"], s["has_flag"])
        _, rr, rd = ref.step(a.tolist())
        orr, od = orc.step(a)
        assert_state_equal(orc.state(), rs.snapshot(ref, ce.cfg.hp_scale), f"seed {seed} t={t}", float_hp=bool(ce.cfg.hp_float))
        assert rd == od
        assert np.array_equal(bits(np.array(rr, dtype=np.float32)), bits(orr)), (seed, t, rr, orr)
        if t % 5 == 0:
            ro, rm = rs.observations(ref)
            oo, om = orc.observe()
            assert np.array_equal(ro, oo), (seed, t)
            assert np.array_equal(bits(rm), bits(om)), (seed, t, rm, om)
    assert np.array_equal(rs.agent_metrics(ref), orc.state()["stats"])


def test_metrics_dict_feeds_the_reference_metrics_logger():
    """env.metrics rebuilt from the agent-level counters (metrics_dict) is harvested by the reference's
    MetricsLogger (metrics_logger.py:137-159) exactly like the reference env's own dict."""
    import importlib

    from marl_ctf_development_b200.env import metrics_dict

    rs.reference_modules()
    ml = importlib.import_module("metrics_logger")
    ec = rs.experiment_env_config("8_arena")
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=5, env_id=5)
    orc = OracleEnv(ce, seed=5, env_id=5)
    pol = traces.make_policy("seek", ce)
    rng = np.random.default_rng(2)
    for t in range(300):
        s = rs.snapshot(ref, ce.cfg.hp_scale)
        a = pol(rng, s["pos"], s["has_flag"])
        ref.step(a.tolist())
        orc.step(a)
    ours = metrics_dict(ce, orc.state()["stats"], orc.state()["visits"])
    team_types = {t: sorted({ce.AGENT_TYPES[i] for i in range(ce.N_AGENTS) if ce.AGENT_TEAMS[i] == t}) for t in (0, 1)}
    loggers = []
    for metrics in (ref.metrics, ours):
        lg = ml.MetricsLogger(1, 0, 0, team_types, list(range(ce.N_AGENTS)), symmetric_teams=False)
        lg.harvest_metrics(metrics, "t0_m0", 0, 0.5, team_idx=0)
        lg.harvest_metrics(metrics, "t1_m0", 0, 0.25, team_idx=1)
        loggers.append(lg)
    a, b = loggers
    for label in ("t0_m0", "t1_m0"):
        for metric, val in a.metrics[label].items():
            if isinstance(val, list):
                assert val == b.metrics[label][metric], (label, metric)
            else:
                assert {k: v for k, v in val.items()} == {k: v for k, v in b.metrics[label][metric].items()}, (label, metric)
    assert a.metrics["t0_m0"]["team_steps_adj_teammate"][0] > 0  # something was actually harvested


def test_multi_lethal_turns_and_carrier_kills_match_the_reference():
    """One-hit-kill settings for 4 episodes: several lethal hits inside one actor's turn (chained respawns whose
    windows depend on each other) and killed flag carriers must have happened, and match step by step."""
    name, exp, overrides, kind = [c for c in KWARG_CASES if c[0] == "one_hit_kills_arena"][0]
    ec = build_env_config(rs.experiment_env_config(exp), overrides)
    ec["GAME_STEPS"] = 400
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=606, env_id=9)
    orc = OracleEnv(ce, seed=606, env_id=9)
    rng = np.random.default_rng(4)
    multi_lethal, carrier_kills = 0, 0
    for episode in range(4):
        if episode:
            ref.reset()
            orc.reset()
        prev = np.zeros((13, ce.N_AGENTS), dtype=np.int64)
        for t in range(400):
            s = rs.snapshot(ref, ce.cfg.hp_scale)
            a = traces.seek_actions_batch(rng, ce, s["pos"][None], s["has_flag"][None], eps=0.15)[0]
            _, rr, rd = ref.step(a.tolist())
            orr, od = orc.step(a)
            st = orc.state()
            assert_state_equal(st, rs.snapshot(ref, ce.cfg.hp_scale), f"ep{episode} t={t}")
            assert np.array_equal(bits(np.array(rr, dtype=np.float32)), bits(orr)) and rd == od
            cur = st["stats"]
            multi_lethal += int(((cur[1] - prev[1]) >= 2).sum())
            carrier_kills += int((cur[4] - prev[4]).sum())
            prev = cur
            if t % 20 == 0:
                ro, rm = rs.observations(ref)
                oo, om = orc.observe()
                assert np.array_equal(ro, oo) and np.array_equal(bits(rm), bits(om))
        assert np.array_equal(rs.agent_metrics(ref), orc.state()["stats"])
    assert multi_lethal >= 2 and carrier_kills >= 10, (multi_lethal, carrier_kills)


def test_visitation_maps_wrap_at_256_like_numpy_uint8():
    """Agents that stand still for 300 steps: the reference's uint8 visitation maps overflow and wrap (gridworld_ctf.py:469, :486)."""
    ec = rs.experiment_env_config("0_the_split")
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=1, env_id=1)
    orc = OracleEnv(ce, seed=1, env_id=1)
    a = np.full(ce.N_AGENTS, 4, dtype=np.uint8)
    for t in range(300):
        ref.step(a.tolist())
        orc.step(a)
    vm = np.stack([ref.metrics["agent_visitation_maps"][i] for i in range(ce.N_AGENTS)])
    assert np.array_equal(vm, orc.state()["visits"])
    assert vm.max() == (301 % 256)  # 1 at reset + 300 steps, wrapped


@pytest.mark.parametrize("exp", rs.EXPERIMENTS)
def test_full_episode_with_the_aggressive_policy(exp):
    """Every experiment config for one whole episode (+2 steps past done) under the flag-seeking / second-action-heavy
    policy the GPU soak tests use: captures, terminal margins, mining, placing and respawns all occur."""
    ec = rs.experiment_env_config(exp)
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=808, env_id=3)
    orc = OracleEnv(ce, seed=808, env_id=3)
    rng = np.random.default_rng(8)
    for t in range(ec["GAME_STEPS"] + 2):
        s = rs.snapshot(ref, ce.cfg.hp_scale)
        a = traces.seek_actions_batch(rng, ce, s["pos"][None], s["has_flag"][None], eps=0.35, second_p=0.5)[0]
        _, rr, rd = ref.step(a.tolist())
        orr, od = orc.step(a)
        assert_state_equal(orc.state(), rs.snapshot(ref, ce.cfg.hp_scale), f"{exp} t={t}")
        assert np.array_equal(bits(np.array(rr, dtype=np.float32)), bits(orr)) and rd == od, (exp, t, rr, orr)
        if t % 25 == 0 or t >= ec["GAME_STEPS"] - 1:
            ro, rm = rs.observations(ref)
            oo, om = orc.observe()
            assert np.array_equal(ro, oo) and np.array_equal(bits(rm), bits(om)), (exp, t)
    st = orc.state()
    assert np.array_equal(rs.agent_metrics(ref), st["stats"])
    assert np.array_equal(np.stack([ref.metrics["agent_visitation_maps"][i] for i in range(ce.N_AGENTS)]), st["visits"])


@pytest.mark.parametrize("name", rs.alt_experiment_names() or ["<no alt_exp in this tree>"])
def test_alternative_experiment_configs_also_match(name):
    """alt_exp/*.py (arena, arena_ii, jailbreak_ii, ... maps that the nine main scripts do not use): every scenario dict
    of scenarios.py that has a script goes through the config compiler and the oracle, step by step against the reference."""
    if name.startswith("<"):
        pytest.skip("alt_exp scripts need the reference source tree")
    ec = rs.alt_experiment_env_config(name)
    ce = compile_config(**ec)
    ref = rs.make_injected_env(ec, seed=7, env_id=77)
    orc = OracleEnv(ce, seed=7, env_id=77)
    assert env_dims(ce) == ref.get_env_dims()
    assert [int(t) for t in ce.TILES_USED] == [int(t) for t in ref.TILES_USED]
    pol = traces.make_policy("builder" if 3 in ce.AGENT_TYPES.values() else "seek", ce)
    rng = np.random.default_rng(11)
    for t in range(150):
        s0 = rs.snapshot(ref, ce.cfg.hp_scale)
        a = pol(rng, s0["pos"], s0["has_flag"])
        _, rr, rd = ref.step(a.tolist())
        orr, od = orc.step(a)
        assert_state_equal(orc.state(), rs.snapshot(ref, ce.cfg.hp_scale), f"{name} t={t}")
        assert rd == od and np.array_equal(bits(np.array(rr, dtype=np.float32)), bits(orr))
        if t % 10 == 0:
            ro, rm = rs.observations(ref)
            oo, om = orc.observe()
            assert np.array_equal(ro, oo) and np.array_equal(bits(rm), bits(om))
    assert np.array_equal(rs.agent_metrics(ref), orc.state()["stats"])
