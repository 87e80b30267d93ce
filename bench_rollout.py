#!/usr/bin/env python
"""BASELINE config 5: self-play rollout on 8_arena with the GPU env feeding the agent_network-style policy.

    python bench_rollout.py [--envs B] [--steps K] [--warmup W] [--policy-dtype bf16|fp32] [--chunk S]
    torchrun --nproc-per-node N bench_rollout.py ...      (env sharding, one rank per GPU)

Reports, as one JSON line from rank 0, agent-steps/s of the whole loop (policy forward for both teams on the
observation buffers the step kernel wrote, action sampling, env.step) next to the env-only figure.  This is a
functional integration of SURVEY §8f row N1: the policy is stock torch/cuDNN (≈3.9 MFLOP per agent-step) and
dominates the time; it is outside the hot path this repo accelerates, so this number is not the headline.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config
    from marl_ctf_development_b200.policy import CtfPolicy

    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--policy-dtype", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--chunk", type=int, default=32768, help="policy samples per forward")
    ap.add_argument("--channels-last", action="store_true", help="run the convolutions in NHWC (stock torch option)")
    ap.add_argument("--obs-dtype", choices=["float32", "bfloat16"], default=None,
                    help="env observation buffer type (default: bfloat16 when the policy runs in bf16)")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.envs
    obs_dtype = args.obs_dtype or ("bfloat16" if args.policy_dtype == "bf16" else "float32")
    torch.backends.cudnn.benchmark = True
    env = GridworldCtfGPU(**experiment_env_config("8_arena"), num_envs=B, device=dev, seed=0, env_id_base=rank * B,
                          reverse_team1_actions=True, stats="counters", obs_dtype=getattr(torch, obs_dtype))
    N, C, G, M = env.N_AGENTS, env.n_channels, env.GRID_SIZE, env.meta_size
    torch.manual_seed(rank)
    pols = [CtfPolicy(9, C, G, M).to(dev).eval() for _ in range(2)]
    if args.channels_last:
        pols = [p.to(memory_format=torch.channels_last) for p in pols]
    teams = [torch.tensor([i for i in range(N) if env.AGENT_TEAMS[i] == t], device=dev) for t in (0, 1)]
    actions = torch.empty((B, N), dtype=torch.uint8, device=dev)
    amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.policy_dtype == "bf16")

    @torch.no_grad()
    def policy_step(obs, meta):
        for idx, pol in zip(teams, pols):
            k = idx.numel()
            g = obs[:, idx].reshape(B * k, C, G, G)
            if args.channels_last:
                g = g.contiguous(memory_format=torch.channels_last)
            m = meta[:, idx].reshape(B * k, M)
            f = env.use_action_mask[idx].unsqueeze(0).expand(B, k).reshape(B * k)
            outs = []
            for s in range(0, B * k, args.chunk):
                with amp:
                    outs.append(pol.get_action(g[s : s + args.chunk], m[s : s + args.chunk], f[s : s + args.chunk]))
            actions[:, idx] = torch.cat(outs).reshape(B, k).to(torch.uint8)

    def run(k_steps, with_policy):
        obs, meta = env.obs, env.meta
        for _ in range(k_steps):
            if with_policy:
                policy_step(obs, meta)
            obs, meta, _, dones, _ = env.step(actions)

    def timed(k_steps, with_policy):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(k_steps, with_policy)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    env.reset()
    actions.fill_(4)
    run(args.warmup, True)
    ms_full = timed(args.steps, True)
    ms_env = timed(args.steps, False)
    if rank == 0:
        total = world * B * N * args.steps
        print(json.dumps({
            "metric": "rollout_agent_steps_per_sec", "value": total / (ms_full * 1e-3), "unit": "agent-steps/s",
            "env_only_value": total / (ms_env * 1e-3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_full / args.steps, "env_ms_per_step": ms_env / args.steps,
            "config": {"workload": f"8_arena self-play rollout, B={B} envs/GPU, CtfPolicy (agent_network.py architecture) "
                                   f"for both teams, policy dtype {args.policy_dtype}, {obs_dtype} observations"
                                   f"{', channels_last' if args.channels_last else ''}", "envs_per_gpu": B},
            "data": "synthetic (random-init policies)",
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
