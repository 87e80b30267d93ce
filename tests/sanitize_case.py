"""Small deterministic workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).

    python tests/sanitize_case.py && compute-sanitizer --tool racecheck python tests/sanitize_case.py

Covers reset, step (all obs dtypes, statistics + visitation maps), observe with explicit flags, stats_sum and
the host-buffer step on three scenarios; checks the final state against the oracle so a silent corruption fails.
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # traces.py

import numpy as np
import torch

import traces
from marl_ctf_development_b200 import GridworldCtfGPU, compile_config, experiment_env_config
from oracle.ctf_oracle import OracleBatch


def main():
    for exp, dtype in (("8_arena", torch.float32), ("7_gridlocked", torch.uint8), ("0_the_split", torch.bfloat16)):
        ec = experiment_env_config(exp)
        B = 7
        env = GridworldCtfGPU(**ec, num_envs=B, device="cuda:0", seed=3, stats="full", obs_dtype=dtype)
        orc = OracleBatch(compile_config(**ec), B, seed=3)
        pol = traces.make_policy("builder", env.ce)
        rng = np.random.default_rng(0)
        a_h = torch.empty((B, env.N_AGENTS), dtype=torch.uint8).pin_memory()
        r_h = torch.empty((B, env.N_AGENTS), dtype=torch.float32).pin_memory()
        d_h = torch.empty((B,), dtype=torch.uint8).pin_memory()
        for t in range(40):
            st = orc.state()
            a = np.stack([pol(rng, st["pos"][b], st["has_flag"][b]) for b in range(B)])
            if t % 2:
                a_h.copy_(torch.from_numpy(a))
                env.step_host(a_h, r_h, d_h)
            else:
                env.step(torch.from_numpy(a).cuda())
            orc.step(a)
        env.observe(reverse_flags=[1] * env.N_AGENTS, into_new=True)
        env.stats_sum(all_reduce=False)
        torch.cuda.synchronize()
        sg, so = env.get_state(), orc.state()
        for k in ("grid", "pos", "hp_q", "has_flag", "inventory", "captures", "stats", "visits"):
            assert np.array_equal(np.asarray(sg[k]).astype(np.int64), np.asarray(so[k]).astype(np.int64)), (exp, k)
        env.reset()
        env.close()
    print("sanitize_case ok")


if __name__ == "__main__":
    main()
