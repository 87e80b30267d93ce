"""Action-trace generators shared by the parity tests and tests/golden/make_golden.py.

Uniform random actions rarely capture flags, so next to them there
is a 'seek' policy that walks agents at the flags, and a 'builder' mix that makes
miners mine / place and vaulters jump.  Policies only read positions / has_flag,
which every implementation under test exposes, and are deterministic given rng.
"""
from __future__ import annotations

import numpy as np


def uniform_actions(rng: np.random.Generator, n_agents: int) -> np.ndarray:
    return rng.integers(0, 9, n_agents).astype(np.uint8)


def seek_actions(rng, ce, pos, has_flag, eps=0.25, second_p=0.35) -> np.ndarray:
    """Greedy walk to the opponent flag (own flag when carrying); eps-random otherwise.

    ce: CompiledEnv; pos: [N,2]; has_flag: [N].  Types 2/3 use their second action set
    (vault / place) with probability second_p so those paths fire on the way.
    """
    n = ce.N_AGENTS
    out = np.zeros(n, dtype=np.uint8)
    for i in range(n):
        team, typ = ce.AGENT_TEAMS[i], ce.AGENT_TYPES[i]
        if rng.random() < eps:
            out[i] = rng.integers(0, 9)
            continue
        tr, tc = ce.FLAG_POSITIONS[team] if has_flag[i] else ce.FLAG_POSITIONS[1 - team]
        dr, dc = int(tr) - int(pos[i][0]), int(tc) - int(pos[i][1])
        choices = []
        if dr < 0:
            choices.append(0)
        if dr > 0:
            choices.append(1)
        if dc > 0:
            choices.append(2)
        if dc < 0:
            choices.append(3)
        a = int(rng.choice(choices)) if choices else 4
        if a < 4 and typ in (2, 3) and rng.random() < second_p:
            a += 5
        out[i] = a
    return out


def make_policy(kind: str, ce):
    if kind == "uniform":
        return lambda rng, pos, has_flag: uniform_actions(rng, ce.N_AGENTS)
    if kind == "seek":
        return lambda rng, pos, has_flag: seek_actions(rng, ce, pos, has_flag)
    if kind == "builder":
        return lambda rng, pos, has_flag: seek_actions(rng, ce, pos, has_flag, eps=0.5, second_p=0.6)
    raise ValueError(kind)


OBS_EVERY = 20


def snap_after_step(t: int, env_step: int, game_steps: int) -> bool:
    """Whether the golden traces hold an observation after recorded step t (t counts from 1)."""
    return t % OBS_EVERY == 0 or env_step in (game_steps - 1, game_steps, game_steps + 1)


def seek_actions_batch(rng, ce, pos, has_flag, eps=0.25, second_p=0.35) -> np.ndarray:
    """Vectorised seek policy for a whole batch: pos [B,N,2], has_flag [B,N] -> uint8 [B,N] (soak tests)."""
    pos = np.asarray(pos, dtype=np.int64)
    B, N = pos.shape[0], ce.N_AGENTS
    teams = np.array([ce.AGENT_TEAMS[i] for i in range(N)])
    types = np.array([ce.AGENT_TYPES[i] for i in range(N)])
    flags = np.array([ce.FLAG_POSITIONS[0], ce.FLAG_POSITIONS[1]], dtype=np.int64)
    target = np.where(np.asarray(has_flag, dtype=bool)[..., None], flags[teams][None], flags[1 - teams][None])
    d = target - pos
    # candidate unit moves toward the target: U, D, R, L
    want = np.stack([d[..., 0] < 0, d[..., 0] > 0, d[..., 1] > 0, d[..., 1] < 0], axis=-1)
    score = np.where(want, rng.random((B, N, 4)), -1.0)
    a = np.where(want.any(-1), score.argmax(-1), 4)
    second = (a < 4) & np.isin(types, (2, 3))[None] & (rng.random((B, N)) < second_p)
    a = np.where(second, a + 5, a)
    rnd = rng.random((B, N)) < eps
    a = np.where(rnd, rng.integers(0, 9, (B, N)), a)
    return a.astype(np.uint8)
