"""TEST / BASELINE INFRASTRUCTURE — times the UNMODIFIED Python reference on this machine's host cores
(BASELINE.md §3): one ``GridworldCtf`` per process, seeded random actions, and per env step what the reference's
callers do (ppo.py:66-98, utils.py:528-553): ``standardise_state`` + ``get_env_metadata`` for every agent, then
``step``.  The reference is imported from /root/reference, or from the byte-compiled oracle/_ref where that tree
does not exist (oracle/build_ref.py).

    python oracle/ref_cpu_baseline.py [--experiment 8_arena] [--seconds 10] [--procs N] [--step-only]

Prints ONE JSON line: aggregate agent-steps/s over all processes (sum of steps / slowest process's time).
bench.py runs this in a subprocess for its ``cpu_baseline`` (kind "reference") and for ``--impl reference``.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def worker(args):
    name, w, seconds, with_obs, max_steps = args
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

    def steps(k0, k1):
        for t in range(k0, k1):
            if t % T == 0:
                env.reset()
            if with_obs:
                for i in range(n):
                    env.standardise_state(i, reverse_grid=(env.AGENT_TEAMS[i] != 0))
                    env.get_env_metadata(i)
            env.step(acts[t % T].tolist())

    steps(0, 25)  # warm-up
    done, chunk = 25, 25
    t0 = time.perf_counter()
    start = done
    while time.perf_counter() - t0 < seconds and (max_steps is None or done - start < max_steps):
        steps(done, done + chunk)
        done += chunk
    return time.perf_counter() - t0, (done - start) * n


def measure(experiment="8_arena", seconds=10.0, procs=None, with_obs=True, max_steps=None) -> dict:
    from oracle import ref_shim as rs

    if not rs.available():
        raise RuntimeError("the reference is neither under /root/reference nor byte-compiled in oracle/_ref")
    procs = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(worker, [(experiment, w, seconds, with_obs, max_steps) for w in range(procs)])
    total = sum(r[1] for r in res)
    slowest = max(r[0] for r in res)
    return {
        "experiment": experiment, "processes": procs, "agent_steps": total, "seconds": slowest,
        "agent_steps_per_s": total / slowest, "with_observations": with_obs,
        "reference": "source tree" if rs.source_tree_available() else "oracle/_ref (byte-compiled)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--experiment", default="8_arena")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--step-only", action="store_true")
    ap.add_argument("--max-steps", type=int, default=0, help="stop every process after this many env steps (0: time-bound only)")
    a = ap.parse_args()
    print(json.dumps(measure(a.experiment, a.seconds, a.procs or None, not a.step_only, a.max_steps or None)), flush=True)


if __name__ == "__main__":
    main()
