"""Seeded generator of random but valid scenario dicts (the scenarios.py format, scenarios.py:5-82) and agent
configs, so the drop-in claim is checked on maps the reference's authors never wrote.

Validity rules, each needed by the reference itself:
  * SPAWN_POSITIONS at least 1 cell from every edge (gridworld_ctf.py:773 warning) and their 3x3 windows free of
    static tiles, flags and starting agents' neighbours are fine (agents move) -> respawn always finds k >= 1;
  * flags, starting positions and spawns on distinct cells; static tiles never on flags / starts / spawn windows.
"""
from __future__ import annotations

import numpy as np


class _Retry(Exception):
    pass


DECIMAL_HP_SEEDS = tuple(range(100, 108))   # seeds whose HP tables are decimals (tenths / twentieths): IEEE-double HP mode


def random_env_config(seed: int) -> dict:
    """Deterministic in ``seed``; internally re-draws when a placement cannot be completed.  Seeds in DECIMAL_HP_SEEDS get
    HP / damage / heal / vault quantities that are not dyadic rationals (0.3, 1.15, ...), which no fixed point holds exactly."""
    for attempt in range(1000):
        try:
            ec = _draw(np.random.default_rng([seed, attempt]))
            break
        except _Retry:
            continue
    else:
        raise RuntimeError("no valid scenario found")
    if seed in DECIMAL_HP_SEEDS:
        rng = np.random.default_rng([seed, 12345])
        d = lambda lo, hi: float(rng.integers(lo, hi)) / 20.0  # noqa: E731
        ec.update({
            "AGENT_TYPE_HP": {t: d(11, 120) for t in range(4)},
            "AGENT_TYPE_DAMAGE": {t: d(0, 45) for t in range(4)},
            "AGENT_HP_HEALING_PER_STEP": d(0, 12),
            "GUARDIAN_DAMAGE_MULTIPLIER": float(rng.choice([1.0, 1.7, 2.3, 5.0])),
            "VAULT_HP_COST": d(0, 40),
            "VAULT_MIN_HP": d(0, 60),
        })
    return ec


def _draw(rng) -> dict:
    seed = int(rng.integers(0, 1 << 30))
    G = int(rng.integers(7, 17))
    n_agents = int(rng.integers(2, 9))
    flip = [None, 0, 1, 2][int(rng.integers(0, 4))]

    def cell(lo=0, hi=None):
        hi = G if hi is None else hi
        return (int(rng.integers(lo, hi)), int(rng.integers(lo, hi)))

    taken = set()
    spawns = {}
    for t in (0, 1):
        for _try in range(200):
            c = cell(1, G - 1)
            win = {(c[0] + dr, c[1] + dc) for dr in (-1, 0, 1) for dc in (-1, 0, 1)}
            if not (win & taken):
                spawns[t] = c
                taken |= win
                break
        else:
            raise _Retry
    reserved = set(taken)  # spawn windows stay free of static tiles and flags

    def fresh():
        for _try in range(500):
            c = cell()
            if c not in taken:
                taken.add(c)
                return c
        raise _Retry

    flags = {0: fresh(), 1: fresh()}
    captures = {0: flags[0], 1: flags[1]} if rng.random() < 0.7 else {0: fresh(), 1: fresh()}
    starts = {}
    for i in range(8):
        # start near the own spawn most of the time (like the shipped maps), anywhere otherwise
        for _try in range(500):
            if rng.random() < 0.7:
                s = spawns[i % 2]
                c = (s[0] + int(rng.integers(-1, 2)), s[1] + int(rng.integers(-1, 2)))
            else:
                c = cell()
            if c not in starts.values() and c not in flags.values() and 0 <= c[0] < G and 0 <= c[1] < G:
                starts[i] = c
                break
        else:
            raise _Retry
    taken |= set(starts.values())
    blocks, destr = [], []
    for _ in range(int(rng.integers(0, G * G // 5))):
        c = cell()
        if c in taken or c in reserved:
            continue
        taken.add(c)
        (blocks if rng.random() < 0.5 else destr).append(c)
    # a slice entry like the shipped maps use, clipped by numpy where it runs past the grid
    if rng.random() < 0.5:
        r = int(rng.integers(0, G))
        c0 = int(rng.integers(0, G - 1))
        cells = {(r, c) for c in range(c0, min(c0 + 3, G))}
        if not (cells & (taken | reserved)):
            taken |= cells
            (blocks if rng.random() < 0.5 else destr).append((r, slice(c0, c0 + 3)))
    scenario = {
        "SCENARIO_NAME": f"random-{seed}", "GRID_SIZE": G, "FLIP_AXIS": flip,
        "FLAG_POSITIONS": flags, "CAPTURE_POSITIONS": captures, "SPAWN_POSITIONS": spawns,
        "AGENT_STARTING_POSITIONS": starts, "BLOCK_TILE_SLICES": blocks, "DESTRUCTIBLE_TILE_SLICES": destr,
    }
    types = [int(rng.integers(0, 4)) for _ in range(n_agents)]
    # get_env_metadata reads agent_hp[type] (gridworld_ctf.py:1041): types must be valid agent ids
    types = [min(t, n_agents - 1) for t in types]
    if rng.random() < 0.6:
        teams = [i % 2 for i in range(n_agents)]
    else:
        teams = [int(rng.integers(0, 2)) for _ in range(n_agents)]
        teams[0], teams[1] = 0, 1
    q = lambda lo, hi: float(rng.integers(lo, hi)) / 4.0  # noqa: E731
    return {
        "GRID_SIZE": G,
        "AGENT_CONFIG": {i: {"team": teams[i], "type": types[i]} for i in range(n_agents)},
        "SCENARIO": scenario,
        "GAME_STEPS": int(rng.integers(30, 120)),
        "USE_ADJUSTED_REWARDS": bool(rng.random() < 0.5),
        "HOME_FLAG_CAPTURE": bool(rng.random() < 0.3),
        "DROP_FLAG_WHEN_NO_HP": bool(rng.random() < 0.2),
        "MAP_SYMMETRY_CHECK": False,
        "AGENT_TYPE_HP": {t: q(4, 44) for t in range(4)},
        "AGENT_TYPE_DAMAGE": {t: q(0, 9) for t in range(4)},
        "AGENT_HP_HEALING_PER_STEP": q(0, 5),
        "TAG_PROBABILITY": float(rng.choice([0.25, 0.5, 0.75, 0.9, 1.0])),
        "GUARDIAN_DAMAGE_MULTIPLIER": float(rng.choice([1.0, 2.0, 5.0])),
        "VAULT_HP_COST": q(0, 8),
        "VAULT_MIN_HP": q(0, 12),
    }
