"""JSON episode traces for the reference's three.js viewer.

Same wire format as ``utils.duel_json`` (utils.py:728-814): ``grid_size, flag_pos, spawn_pos, agent_config,
block_tiles, destructible_tiles`` from the reset state, then per step ``movement`` (position deltas + has_flag),
``tiles`` (destructible tiles, type 0 = intact / 1 = damaged) and ``scores``.  One env of the batch is exported
while the whole batch is stepped on the GPU.
"""
from __future__ import annotations

import json

import numpy as np
import torch

from .rollout import _require_dense_obs, batched_action


def _tiles(grid: np.ndarray) -> list:
    out = [{"x": int(x), "z": int(z), "type": 0} for z, x in zip(*np.where(grid == 2))]
    out += [{"x": int(x), "z": int(z), "type": 1} for z, x in zip(*np.where(grid == 3))]
    return out


@torch.no_grad()
def duel_json(env, agent, opponent, env_index=0, max_steps=256, fname=None) -> dict:
    """Plays one duel (utils.py:728-814) on every env of ``env`` and records env ``env_index``."""
    if not env.ce.cfg.reverse_team1_actions:
        raise ValueError("create the env with reverse_team1_actions=True")
    _require_dense_obs(env)
    B, N, dev = env.num_envs, env.N_AGENTS, env.device
    C, G, M = env.n_channels, env.GRID_SIZE, env.meta_size
    obs, meta, _ = env.reset()
    st = env.get_state(env_index=env_index)   # only the recorded env leaves the device
    grid0 = st["grid"][0]
    out = {
        "grid_size": int(env.GRID_SIZE),
        "flag_pos": {f"{k}": {"x": int(v[1]), "z": int(v[0])} for k, v in env.FLAG_POSITIONS.items()},
        "spawn_pos": {f"{k}": {"x": int(v[1]), "z": int(v[0])} for k, v in env.SPAWN_POSITIONS.items()},
        "agent_config": [
            {
                "team": int(env.AGENT_TEAMS[i]), "type": int(env.AGENT_TYPES[i]),
                "start_x": int(env.AGENT_STARTING_POSITIONS[i][1]), "start_z": int(env.AGENT_STARTING_POSITIONS[i][0]),
            }
            for i in range(N)
        ],
        "block_tiles": [{"x": int(x), "z": int(z)} for z, x in zip(*np.where(grid0 == 1))],
        "destructible_tiles": _tiles(grid0),
    }
    teams = [torch.tensor([i for i in range(N) if env.AGENT_TEAMS[i] == t], device=dev) for t in (0, 1)]
    actions = torch.empty((B, N), dtype=torch.uint8, device=dev)
    movement, tiles, scores = [], [], []
    pos = st["pos"][0].astype(np.int64)
    step_count = 0
    while True:
        step_count += 1
        for idx, pol in zip(teams, (agent, opponent)):
            k = idx.numel()
            if k:
                a = batched_action(
                    pol, obs[:, idx].reshape(B * k, C, G, G).float(), meta[:, idx].reshape(B * k, M),
                    env.use_action_mask[idx].unsqueeze(0).expand(B, k).reshape(B * k),
                )
                actions[:, idx] = a.reshape(B, k).to(torch.uint8)
        obs, meta, _, _, _ = env.step(actions)
        st = env.get_state(env_index=env_index)
        new_pos = st["pos"][0].astype(np.int64)
        flags = st["has_flag"][0]
        movement.append(
            [{"x": int(new_pos[i, 1] - pos[i, 1]), "z": int(new_pos[i, 0] - pos[i, 0]), "has_flag": int(flags[i])} for i in range(N)]
        )
        tiles.append(_tiles(st["grid"][0]))
        caps = st["captures"][0]
        scores.append([{"t0": int(caps[0]), "t1": int(caps[1])}])
        pos = new_pos
        if step_count > max_steps or step_count >= env.GAME_STEPS:
            break
    out["movement"], out["tiles"], out["scores"] = movement, tiles, scores
    if fname is not None:
        with open(fname, "w") as f:
            json.dump(out, f, indent=4)
    return out
