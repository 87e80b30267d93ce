"""The CPU oracle replays the committed reference traces bit for bit (tests/golden/, made by make_golden.py)."""
import numpy as np
import pytest

import traces
from helpers import STATE_KEYS, assert_state_equal, bits, compiled, golden_ids, golden_traces
from oracle.ctf_oracle import OracleEnv, f64_to_f16_to_f32


@pytest.mark.parametrize("exp,kind,path", golden_traces(), ids=golden_ids())
def test_oracle_replays_reference_trace(exp, kind, path):
    tr = np.load(path)
    ce = compiled(exp)
    env = OracleEnv(ce, seed=int(tr["seed"]), env_id=int(tr["env_id"]))
    t, oi = 0, 0  # recorded steps so far, next observation snapshot
    for episode, steps in enumerate(tr["episode_lengths"]):
        if episode:
            env.reset()
        assert tr["obs_steps"][oi] == t
        _check_obs(env, tr, oi, f"{exp}/{kind} reset obs ep{episode}")
        oi += 1
        for _ in range(int(steps)):
            r, d = env.step(tr["actions"][t])
            st = env.state()
            assert_state_equal(st, {k: tr[k][t] for k in STATE_KEYS}, f"{exp}/{kind} t={t}")
            assert np.array_equal(bits(r), bits(tr["rewards"][t])), (t, r, tr["rewards"][t])
            assert d == bool(tr["done"][t])
            assert st["episode"] == episode
            t += 1
            if traces.snap_after_step(t, st["step"], ce.GAME_STEPS):
                assert tr["obs_steps"][oi] == t
                _check_obs(env, tr, oi, f"{exp}/{kind} t={t}")
                oi += 1
        st = env.state()
        assert np.array_equal(st["stats"], tr["stats"][episode]), f"{exp}/{kind} stats ep{episode}"
        assert np.array_equal(st["visits"], tr["visits"][episode]), f"{exp}/{kind} visits ep{episode}"
    assert oi == len(tr["obs_steps"])


def _check_obs(env, tr, i, where):
    obs, meta = env.observe()
    assert np.array_equal(obs, tr["obs"][i].astype(np.float32)), where + " obs"
    assert np.array_equal(bits(meta), bits(tr["meta"][i])), where + " meta"
    obs8, _ = env.observe(u8=True)
    assert np.array_equal(obs8, tr["obs"][i]), where + " obs u8"
    obs_f, meta_f = env.observe_fast()  # the CPU-baseline writer gives the same bytes
    assert np.array_equal(obs_f, obs) and np.array_equal(bits(meta_f), bits(meta)), where + " fast writer"


def test_half_rounding_matches_numpy():
    rng = np.random.default_rng(0)
    xs = np.concatenate(
        [
            rng.random(20000) * 4.0,
            np.arange(0, 1200) / 500.0,                       # step / GAME_STEPS
            np.array([(a + 1) / (b + 1) for a in range(70) for b in range(70)]),  # capture ratios
            10.0 ** rng.uniform(-9, 5.2, 5000),
            np.array([0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322387695312e-08, 6.1e-5]),
        ]
    )
    with np.errstate(over="ignore"):
        want = xs.astype(np.float16).astype(np.float32)
    got = np.array([f64_to_f16_to_f32(float(x)) for x in xs], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
