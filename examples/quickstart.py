"""Quick start: 4 096 arena envs on one GPU, random actions, episode statistics in the reference's schema.

    python examples/quickstart.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config

env_config = experiment_env_config("8_arena")          # == TrainingConfig().env_config of the reference's 8_arena.py
env = GridworldCtfGPU(**env_config, num_envs=4096, device="cuda:0", seed=42, stats="counters")
obs, meta, mask = env.reset()
print("obs", tuple(obs.shape), obs.dtype, "| meta", tuple(meta.shape), "| action mask", tuple(mask.shape))

for t in range(env.GAME_STEPS):
    actions = torch.randint(0, 9, (env.num_envs, env.N_AGENTS), dtype=torch.uint8, device=env.device)
    obs, meta, rewards, dones, mask = env.step(actions)
print("all done:", bool(dones.all()), "| mean terminal reward per agent:", rewards.mean(0).tolist())

metrics = env.episode_stats()                            # env.metrics of the reference, summed over the batch
print("team flag captures:", metrics["team_flag_captures"], "| team tags:", metrics["team_tag_count"])
print("flag captures by agent:", dict(metrics["agent_flag_captures"]))
