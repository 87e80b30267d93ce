#!/usr/bin/env python
"""Throughput of the GridworldCtf step path: agent-steps/s for 8_arena at B envs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs B] [--impl ours|reference]

One "step" is one pass of the hot path over one batch: env.step(actions) for B envs — the fused
kernel advances all N agents of every env and writes their float32 observations + metadata into the
policy input buffers — so a step is B*N agent-steps.  Under torchrun (N > 1) every rank owns B envs
with global ids rank*B .. rank*B+B-1 (weak scaling, no collective inside the step); episode
statistics are all-reduced over NCCL once per episode.  Rank 0 prints ONE JSON line.

--impl reference times the CPU oracle (oracle/ctf_oracle.c, a port of the reference's algorithm; the
reference itself is pure Python and does not travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EXPERIMENT = "8_arena"  # BASELINE.json headline; --experiment selects another config for side measurements
METRIC = "agent_steps_per_sec"
UNIT = "agent-steps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def algorithmic_bytes_per_agent_step(G: int, N: int, C: int, obs_elem_bytes: int = 4) -> float:
    """obs + metadata + reward + (state read+write + actions + done + mask) / N."""
    M = 6 + 2 * N
    state_rw = 2 * (G * G + 2 * N + 2 * N + N + 2 * N + 8)
    return C * G * G * obs_elem_bytes + M * 4 + 4 + (state_rw + N + 1 + N) / N


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = get_reasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(self.period)
        except Exception as exc:  # NVML missing: report it instead of inventing numbers
            self.error = repr(exc)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2)
        out = {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }
        if self.error:
            out["error"] = self.error
        return out


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per k_step launch from the committed ncu capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


# -----------------------------------------------------------------------------------------------------
# CPU legs (oracle port): the one place outside tests/ where oracle/ is executed, as a reported baseline
# -----------------------------------------------------------------------------------------------------
def cpu_leg(envs_per_thread: int, steps: int, warm_steps: int = 20):
    from marl_ctf_development_b200 import compile_config, experiment_env_config
    from oracle.ctf_oracle import OracleBatch

    threads = host_threads()
    ce = compile_config(**experiment_env_config(EXPERIMENT))
    cpu_leg.dims = (ce.GRID_SIZE, ce.N_AGENTS, ce.n_channels)
    n_envs = envs_per_thread * threads
    batch = OracleBatch(ce, n_envs, seed=1)
    batch.run(warm_steps, 1, 1, threads)
    return batch, ce, threads, n_envs


def cpu_baseline_sample(n_agents: int, seconds: float = 8.0) -> dict:
    """Bounded samples (~2 x 8 s) of the same workload on all host threads: the literal port of the reference's
    algorithm (the reported baseline) and, for context, the same port with a tuned observation writer."""
    out = {}
    for key, mode in (("value", 1), ("tuned_value", 2)):
        batch, ce, threads, n_envs = cpu_leg(envs_per_thread=64, steps=0, warm_steps=5)
        steps, chunk = 0, 10
        t1 = time.perf_counter()
        while time.perf_counter() - t1 < seconds:
            batch.run(chunk, 7 + steps, mode, threads)
            steps += chunk
        dt = time.perf_counter() - t1
        out[key] = n_envs * n_agents * steps / dt
        out[key + "_sample"] = f"{n_envs} envs x {steps} steps in {dt:.1f} s"
    return {
        "value": out["value"], "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"oracle/ctf_oracle.c (line-by-line port of gridworld_ctf.py) on {threads} threads, {EXPERIMENT}, per step "
                  f"obs+meta for all agents then step(): {out['value_sample']}",
        "tuned_value": out["tuned_value"],
        "tuned_sample": f"same port with a scatter-style observation writer (not the reference's algorithm): {out['tuned_value_sample']}",
        "python_reference_note": "the unmodified Python reference measured 7.5e3 agent-steps/s per core (BASELINE.md §2); it cannot travel to this box",
    }


def run_reference_arm(args) -> dict:
    """One 'step' = one pass (observations + metadata for all agents, then step) over a bounded batch on all host threads."""
    batch, ce, threads, n_envs = cpu_leg(envs_per_thread=256, steps=0, warm_steps=2)
    for w in range(args.warmup):
        batch.run(1, 2 + w, 1, threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        batch.run(1, 100 + k, 1, threads)
    dt = time.perf_counter() - t0
    value = n_envs * ce.N_AGENTS * args.steps / dt
    sample = (f"oracle/ctf_oracle.c (line-by-line port of gridworld_ctf.py; the Python reference cannot travel): {n_envs} envs "
              f"({threads} threads x 256) x {args.steps} steps of {EXPERIMENT}, obs+meta for all agents then step()")
    return {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(n_envs, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def workload_config(envs_per_gpu: int, n_gpus: int) -> dict:
    from marl_ctf_development_b200 import compile_config, experiment_env_config

    ce = compile_config(**experiment_env_config(EXPERIMENT))
    G, N, C = ce.GRID_SIZE, ce.N_AGENTS, ce.n_channels
    out_mb = envs_per_gpu * N * (C * G * G + 6 + 2 * N) * 4 / 1e6
    return {
        "workload": f"{EXPERIMENT} ({ce.SCENARIO_NAME}, {G}x{G}, {N} agents), B={envs_per_gpu} envs/GPU, "
                    f"uniform random actions 0..8, float32 observations [B,{N},{C},{G},{G}] + metadata [B,{N},{6 + 2 * N}]",
        "envs_per_gpu": envs_per_gpu,
        "n_agents": N,
        "obs_dtype": "float32",
        "parallelism": f"env-sharded x{n_gpus} (no collective in step; NCCL all-reduce of episode stats per episode)",
        "l2": f"per-step output is {out_mb:.0f} MB " + ("(exceeds the 126 MB L2, no explicit flush)" if out_mb > 126 else
              "(fits the 126 MB L2: flushed between timed steps is NOT done, treat as L2-resident side measurement)"),
    }


# -----------------------------------------------------------------------------------------------------
def run_ours(args) -> dict | None:
    import torch
    import torch.distributed as dist

    from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    B, K, W = args.envs, args.steps, args.warmup
    ec = experiment_env_config(EXPERIMENT)
    env = GridworldCtfGPU(**ec, num_envs=B, device=dev, seed=args.seed, env_id_base=rank * B,
                          stats="none" if args.no_stats else "counters",
                          obs_dtype=getattr(torch, args.obs_dtype), packed_obs=args.packed or args.no_dense,
                          dense_obs=not args.no_dense)
    N, G, C = env.N_AGENTS, env.GRID_SIZE, env.n_channels
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 8  # distinct pre-generated action tensors, resident in HBM before the timed region
    actions = torch.randint(0, 9, (n_act, B, N), dtype=torch.uint8, device=dev, generator=gen)
    stats_total = torch.zeros((13, N), dtype=torch.int64, device=dev)

    launches = 0

    def one_step(k):
        nonlocal launches
        env.step(actions[k % n_act])
        launches += 1
        if (k + 1) % env.GAME_STEPS == 0:  # episode over: reduce its statistics, start the next one
            if not args.no_stats:
                stats_total.add_(env.stats_sum(all_reduce=world > 1))
                launches += 1
            env.reset()
            launches += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    step_idx = 0
    for _ in range(W):
        one_step(step_idx)
        step_idx += 1
    graph = None
    if args.graph_steps > 0:
        graph = env.make_step_graph(actions[0], steps_per_replay=args.graph_steps)
    # ---- timed region: K steps, CUDA events on the launching stream, barrier + synchronize on both sides
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if graph is None:
        for _ in range(K):
            one_step(step_idx)
            step_idx += 1
    else:
        for _ in range(K // args.graph_steps):
            graph.replay()
        launches = (K // args.graph_steps) * args.graph_steps
        K = launches
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    gpu_launches = launches

    # ---- e2e: the same steps through the host-buffer entry point (ctf_step_host): H2D actions from pinned
    # memory and D2H rewards + dones inside the timed region; observations stay in the policy's device buffer
    a_host = [actions[i].cpu().pin_memory() for i in range(n_act)]
    r_host = torch.empty((B, N), dtype=torch.float32).pin_memory()
    d_host = torch.empty((B,), dtype=torch.uint8).pin_memory()
    Ke = max(1, min(K, 200))
    for i in range(3):
        env.step_host(a_host[i % n_act], r_host, d_host)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        env.step_host(a_host[i % n_act], r_host, d_host)
        if bool(d_host[0]):
            env.reset()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(times[0]), float(times[1])

    result = None
    if rank == 0:
        agent_steps = world * B * N * K
        value = agent_steps / (ms * 1e-3)
        per_agent_step = algorithmic_bytes_per_agent_step(G, N, C, {"float32": 4, "uint8": 1}.get(args.obs_dtype, 2))
        if args.no_dense:
            per_agent_step -= C * G * G * {"float32": 4, "uint8": 1}.get(args.obs_dtype, 2)
        if args.packed or args.no_dense:
            per_agent_step += env.bits_words_per_agent * 4
        launch_s = ms * 1e-3 / K
        achieved = per_agent_step * B * N / launch_s / 1e9
        peak, peak_src = measured_hbm_peak()
        traffic = ncu_traffic_per_launch()
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": dict(workload_config(B, world), obs_dtype=args.obs_dtype,
                           observation_outputs=("packed only" if args.no_dense else "dense + packed" if args.packed else "dense")),
            "clocks": clocks,
            "e2e": {
                "value": world * B * N * Ke / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": B * N, "d2h_bytes_per_step": B * N * 4 + B,
                "steps": Ke, "api": "GridworldCtfGPU.step_host -> ctf_step_host (pinned host actions in, rewards+dones out, moved over PCIe by the step kernel itself; obs/meta stay in HBM)",
            },
            "gpu_launches": gpu_launches,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic["bytes_per_launch"] if traffic and traffic.get("envs_per_gpu") == B else None,
                "kernel": "k_step<%s,%s>" % ({"float32": "float", "uint8": "uint8_t", "float16": "__half", "bfloat16": "__nv_bfloat16"}[args.obs_dtype], "false" if args.no_stats else "true"),
                "algorithmic_bytes_per_agent_step": per_agent_step,
                "bytes_per_launch": per_agent_step * B * N,
                "launch_ms": launch_s * 1e3,
                "peak_source": peak_src,
            },
            "episode_stats_checksum": int(stats_total.sum().item()),
        }
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline_sample(N)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU (BASELINE.json: 65536)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-stats", action="store_true", help="skip the episode-statistics counters")
    ap.add_argument("--packed", action="store_true", help="side measurement: also write the packed (1 bit/element) observation copy")
    ap.add_argument("--no-dense", action="store_true", help="side measurement: packed observations only (implies --packed)")
    ap.add_argument("--graph-steps", type=int, default=0,
                    help="side measurement: replay a CUDA graph of this many captured steps (launch-bound small batches)")
    ap.add_argument("--experiment", default="8_arena", help="experiment config (side measurements; the headline is 8_arena)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--obs-dtype", choices=["float32", "uint8", "float16", "bfloat16"], default="float32",
                    help="float32 is the drop-in default and the headline; the narrower buffers are reported separately")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    global EXPERIMENT
    EXPERIMENT = args.experiment

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return  # under torchrun only rank 0 runs the CPU arm
        print(json.dumps(run_reference_arm(args)), flush=True)
        return
    result = run_ours(args)
    if result is not None:
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
