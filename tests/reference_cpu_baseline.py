"""Times the UNMODIFIED Python reference on this machine's cores (BASELINE.md §3 plan): one GridworldCtf per
process, seeded random actions, per step standardise_state + get_env_metadata for every agent, then step().

    python tests/reference_cpu_baseline.py [experiment ...]      (needs /root/reference; not run on the GPU box)

Prints one JSON line per experiment.  This is context for the CPU baseline that bench.py reports (the C port): the
reference itself cannot travel to the GPU box.
"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def worker(args):
    name, w, episodes, with_obs = args
    import random
    import warnings

    import numpy as np

    from oracle import ref_shim as rs

    warnings.filterwarnings("ignore")
    ec = rs.experiment_env_config(name)
    env = rs.make_reference_env(ec)
    random.seed(w)
    np.random.seed(w)
    n, T = env.N_AGENTS, ec["GAME_STEPS"]
    acts = np.random.default_rng(w).integers(0, 9, (T, n))

    def episode():
        env.reset()
        for t in range(T):
            if with_obs:
                for i in range(n):
                    env.standardise_state(i, reverse_grid=(env.AGENT_TEAMS[i] != 0))
                    env.get_env_metadata(i)
            env.step(acts[t].tolist())

    episode()  # warm-up
    t0 = time.perf_counter()
    for _ in range(episodes):
        episode()
    return time.perf_counter() - t0, episodes * T * n


def main():
    names = sys.argv[1:] or ["0_the_split", "7_gridlocked", "8_arena"]
    procs = len(os.sched_getaffinity(0))
    for name in names:
        out = {"experiment": name, "processes": procs}
        for with_obs, key in ((True, "full"), (False, "step_only")):
            with mp.Pool(procs) as pool:
                res = pool.map(worker, [(name, w, 3, with_obs) for w in range(procs)])
            out[key + "_agent_steps_per_s"] = sum(r[1] for r in res) / max(r[0] for r in res)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
