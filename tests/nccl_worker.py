"""Worker of tests/test_gpu_multi.py: one rank per GPU, env shard with global ids, NCCL all-reduce of the statistics."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config
from marl_ctf_development_b200.sharding import env_id_base


def main():
    out_dir, B, steps, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    env = GridworldCtfGPU(**experiment_env_config("8_arena"), num_envs=B, device=f"cuda:{local}", seed=seed,
                          env_id_base=env_id_base(B), stats="counters")
    acts = np.random.default_rng(3).integers(0, 9, (steps, world * B, env.N_AGENTS)).astype(np.uint8)
    for t in range(steps):
        env.step(torch.from_numpy(acts[t, rank * B : (rank + 1) * B]).cuda())
    total = env.stats_sum(all_reduce=True)            # NCCL all-reduce over NVLink
    metrics = env.episode_stats()                     # the env.metrics schema from the reduced counters
    st = env.get_state()
    np.save(os.path.join(out_dir, f"total_{rank}.npy"), total.cpu().numpy())
    np.save(os.path.join(out_dir, f"grid_{rank}.npy"), st["grid"])
    with open(os.path.join(out_dir, f"caps_{rank}.json"), "w") as f:
        json.dump({"team_flag_captures": metrics["team_flag_captures"], "team_tag_count": metrics["team_tag_count"]}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
