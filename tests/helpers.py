"""Shared helpers of the parity tests."""
from __future__ import annotations

import glob
import os

import numpy as np

from marl_ctf_development_b200.config import compile_config
from marl_ctf_development_b200.scenario_io import experiment_env_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STATE_KEYS = ("grid", "pos", "hp_q", "has_flag", "inventory", "captures")


def golden_traces():
    """[(experiment, policy, path)] of the committed reference traces."""
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN, "trace_*.npz"))):
        base = os.path.basename(path)[len("trace_") : -len(".npz")]
        exp, kind = base.rsplit("_", 1)
        out.append((exp, kind, path))
    return out


def golden_ids():
    return [f"{e}-{k}" for e, k, _ in golden_traces()]


def compiled(exp: str, **overrides):
    ec = experiment_env_config(exp)
    ec.update(overrides)
    return compile_config(**ec)


def bits(a):
    """fp32 arrays are compared bit for bit."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def hp_bits(a):
    """float64 HP values are compared bit for bit."""
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_state_equal(got: dict, want: dict, where="", float_hp=False):
    """float_hp: the configuration keeps HP as doubles (cfg.hp_float) — 'hp' is compared bitwise instead of 'hp_q'."""
    for k in STATE_KEYS:
        if float_hp and k == "hp_q":
            continue
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), np.asarray(want[k]).astype(np.int64)), (
            f"{where}: {k} differs\n got={got[k]}\nwant={want[k]}"
        )
    if "hp" in got and "hp" in want:   # agent_hp as floats: the oracle tracks it in every mode, the GPU env in float mode
        assert np.array_equal(hp_bits(got["hp"]), hp_bits(want["hp"])), f"{where}: hp differs\n got={got['hp']}\nwant={want['hp']}"
