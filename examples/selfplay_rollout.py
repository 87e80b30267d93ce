"""Self-play rollout + duel + win-rate matrix on GPU envs with agent_network-style policies (BASELINE config 5).

    python examples/selfplay_rollout.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config
from marl_ctf_development_b200.league import winrate_matrix_symmetric
from marl_ctf_development_b200.policy import CtfPolicy
from marl_ctf_development_b200.rollout import batched_duel, collect_rollout

ec = experiment_env_config("8_arena")
env = GridworldCtfGPU(**ec, num_envs=256, device="cuda:0", seed=0, reverse_team1_actions=True, stats="counters", packed_obs=True)
make = lambda: CtfPolicy(9, env.n_channels, env.GRID_SIZE, env.meta_size).cuda()  # noqa: E731
agent, opponent = make(), make()

# one PPO rollout for all 256 envs (ppo.py:31-131), observations stored packed (32x smaller), unpacked per minibatch
ro = collect_rollout(env, agent, opponent, train_team1=True, num_env_steps=100, obs_storage_dtype="packed")
mb = torch.randint(0, ro.grid_states.shape[0], (64,), device="cuda")
print("rollout rows", ro.actions.shape, "| packed obs", tuple(ro.grid_states.shape), "->", tuple(ro.unpack_grid_states(env, mb).shape))

# 256 duels at once (utils.py:500-573)
result = batched_duel(env, agent, opponent, max_steps=256)
print("duel results (+1 win / 0 draw / -1 loss for team 0):", torch.bincount(result + 1, minlength=3).tolist())

# win-rate matrix of a 3-agent league, 32 duels per pair (league_training.py:368-400)
wm = winrate_matrix_symmetric(ec, [agent, opponent, make()], 32, max_steps=128, device="cuda:0")
print({k: round(v, 3) for k, v in wm.items()})
