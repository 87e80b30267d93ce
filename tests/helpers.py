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


def assert_state_equal(got: dict, want: dict, where=""):
    for k in STATE_KEYS:
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), np.asarray(want[k]).astype(np.int64)), (
            f"{where}: {k} differs\n got={got[k]}\nwant={want[k]}"
        )
