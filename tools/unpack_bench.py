"""Bandwidth of ctf_unpack_obs: packed observation blocks (396 B per agent for 8_arena) -> the policy's dtype."""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config

B = 65536
env = GridworldCtfGPU(**experiment_env_config("8_arena"), num_envs=B, device="cuda:0", seed=0, packed_obs=True, dense_obs=False)
gen = torch.Generator(device="cuda").manual_seed(0)
for _ in range(20):
    env.step(torch.randint(0, 9, (B, 8), dtype=torch.uint8, device="cuda", generator=gen))
for dtype in (torch.float32, torch.bfloat16, torch.uint8):
    out = torch.empty((B, 8, env.n_channels, 15, 15), dtype=dtype, device="cuda")
    for _ in range(3):
        env.unpack_obs(env.obs_bits, dtype=dtype, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 20
    for _ in range(reps):
        env.unpack_obs(env.obs_bits, dtype=dtype, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = out.numel() * out.element_size() + env.obs_bits.numel() * 4
    print(json.dumps({"kernel": "k_unpack", "dtype": str(dtype), "agent_blocks": B * 8, "ms": ms, "GBps": nbytes / ms / 1e6,
                      "bytes": nbytes, "frac_of_measured_copy_peak_6549": nbytes / ms / 1e6 / 6549.4}))
