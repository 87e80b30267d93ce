"""A/B sweep of the step kernel's shape on one GPU: warp-per-env kernel vs the persistent warp-specialised kernel
with different logic / stream warp counts (CTF_WS_* are read by ctf_create).  Prints one JSON line per shape.

    python tools/ws_sweep.py [--experiment 8_arena] [--envs 65536] [--steps 200] [--obs-dtype float32] [--shapes quick|full]
"""
import argparse
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config  # noqa: E402


def time_shape(args, env_vars):
    for k in ("CTF_WS", "CTF_WS_LOGIC", "CTF_WS_STREAM", "CTF_WS_CTAS_PER_SM", "CTF_WS_MIN_ENVS", "CTF_K_STEP_CTAS_PER_SM"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env_vars.items()})
    dev = torch.device("cuda", 0)
    env = GridworldCtfGPU(**experiment_env_config(args.experiment), num_envs=args.envs, device=dev, seed=0,
                          stats="none" if args.no_stats else "counters", obs_dtype=getattr(torch, args.obs_dtype),
                          packed_obs=args.packed or args.no_dense, dense_obs=not args.no_dense)
    gen = torch.Generator(device=dev).manual_seed(1)
    acts = torch.randint(0, 9, (8, args.envs, env.N_AGENTS), dtype=torch.uint8, device=dev, generator=gen)
    for i in range(30):
        env.step(acts[i % 8])
    torch.cuda.synchronize()
    best = None
    for rep in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            env.step(acts[i % 8])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        best = ms if best is None else min(best, ms)
    n = env.N_AGENTS
    prof = None
    if hasattr(env._lib, "ctf_debug_ws_profile") and env_vars.get("CTF_WS", 1) != 0:
        import ctypes as C

        buf = (C.c_uint64 * 6)()
        env._lib.ctf_debug_ws_profile(env._handle, buf)        # totals over warm-up + timed steps, all warps
        n_steps = 30 + args.steps * args.reps
        per_env = [v / (n_steps * args.envs) for v in buf]      # cycles per env, summed over the warps of its role
        prof = dict(zip(("fetch", "step", "wait_buffer", "build_publish", "stream_idle", "stream"), (round(x) for x in per_env)))
    env.close()
    del env, acts
    torch.cuda.empty_cache()
    out = {"shape": env_vars, "ms_per_step": round(best, 5), "agent_steps_per_s": round(args.envs * n / best * 1e3)}
    if prof:
        out["cycles_per_env"] = prof
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--experiment", default="8_arena")
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--obs-dtype", default="float32")
    ap.add_argument("--no-stats", action="store_true")
    ap.add_argument("--packed", action="store_true")
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--shapes", default="quick")
    ap.add_argument("--k-step-ctas", type=int, default=-1, help="resident-CTA cap of the warp-per-env kernel for every shape (-1: heuristic)")
    args = ap.parse_args()
    shapes = [{"CTF_WS": 0}]
    if args.shapes == "kstep":       # the warp-per-env kernel only
        shapes = [{"CTF_WS": 0}]
    if args.shapes == "residency":   # resident CTAs per SM of the warp-per-env kernel
        shapes = [{"CTF_WS": 0, "CTF_K_STEP_CTAS_PER_SM": n} for n in (9, 8, 7, 6, 5, 4)]
    if args.shapes in ("residency", "kstep"):
        grid = []
    elif args.shapes == "quick":
        grid = [(16, 8, 1), (12, 8, 1), (20, 8, 1), (24, 8, 1), (16, 4, 1), (8, 4, 2), (12, 4, 2)]
    elif args.shapes == "default":
        grid = []
        shapes.append({"CTF_WS": 1, "CTF_WS_MIN_ENVS": 1})
    elif ";" in args.shapes or "," in args.shapes:
        grid = [tuple(int(x) for x in item.split(",")) for item in args.shapes.split(";") if item]
    else:
        grid = [(l, s, c) for c in (1, 2) for s in (4, 8, 12) for l in (8, 10, 12, 14, 16, 20, 24) if (l + s) <= 32 and (l + s) * c <= 64]
    for l, s, c in grid:
        shapes.append({"CTF_WS": 1, "CTF_WS_LOGIC": l, "CTF_WS_STREAM": s, "CTF_WS_CTAS_PER_SM": c, "CTF_WS_MIN_ENVS": 1})
    for sh in shapes:
        if args.k_step_ctas >= 0:
            sh = dict(sh, CTF_K_STEP_CTAS_PER_SM=args.k_step_ctas)
        try:
            print(json.dumps(dict(time_shape(args, sh), experiment=args.experiment, envs=args.envs, obs_dtype=args.obs_dtype)), flush=True)
        except Exception as exc:  # a shape that does not fit (shared memory) is reported, not fatal
            print(json.dumps({"shape": sh, "error": repr(exc)}), flush=True)


if __name__ == "__main__":
    main()
