"""Config compiler: (AGENT_CONFIG, SCENARIO, ctor kwargs) -> flat ``ctf_config_t``.

Restates what ``GridworldCtf.__init__`` (gridworld_ctf.py:19-350),
``load_scenario`` (:352-381), ``reset`` (:383-477) and ``get_tiles_used``
(:488-499) derive from their arguments, as the POD declared in
include/ctf_b200.h.  The constructor keyword names and defaults are the
reference's, so ``compile_config(**TrainingConfig().env_config)`` works for
every experiment script.
"""
from __future__ import annotations

import ctypes as C
from fractions import Fraction

import numpy as np

MAX_AGENTS = 8
MAX_GRID = 16
MAX_CELLS = 256
N_ACTIONS = 9
N_TYPES = 4
N_METRICS = 13

# agent-level metric families in stats-buffer order (gridworld_ctf.py:456-468)
METRIC_NAMES = (
    "tag_count",
    "respawn_tag_count",
    "flag_pickups",
    "flag_captures",
    "flag_dispossessions",
    "blocks_laid",
    "blocks_mined",
    "blocks_laid_distance_from_own_flag",
    "blocks_laid_distance_from_opp_flag",
    "steps_defending_zone",
    "steps_attacking_zone",
    "steps_adj_teammate",
    "steps_adj_opponent",
)


class CtfConfig(C.Structure):
    """ctypes mirror of ``ctf_config_t`` (include/ctf_b200.h) — keep in the same order."""

    _fields_ = [
        ("tag_threshold", C.c_uint64),
        ("reward_step", C.c_double),
        ("reward_capture", C.c_double),
        ("reward_tag", C.c_double),
        ("capture_punish", C.c_double),
        ("win_margin_scalar", C.c_double),
        ("loss_margin_scalar", C.c_double),
        ("grid_size", C.c_int32),
        ("n_agents", C.c_int32),
        ("n_channels", C.c_int32),
        ("game_steps", C.c_int32),
        ("flip_axis", C.c_int32),
        ("use_adjusted_rewards", C.c_int32),
        ("home_flag_capture", C.c_int32),
        ("drop_flag_when_no_hp", C.c_int32),
        ("hp_scale", C.c_int32),
        ("heal_q", C.c_int32),
        ("vault_cost_q", C.c_int32),
        ("vault_min_q", C.c_int32),
        ("zone_distance", C.c_int32),
        ("guardian_distance", C.c_int32),
        ("tagging_range", C.c_int32),
        ("max_agent_blocks", C.c_int32),
        ("block_pickup_value", C.c_int32),
        ("reverse_team1_actions", C.c_int32),
        ("hp_max_q", C.c_int32 * N_TYPES),
        ("damage_q", C.c_int32 * N_TYPES),
        ("damage_boosted_q", C.c_int32 * N_TYPES),
        ("agent_team", C.c_uint8 * MAX_AGENTS),
        ("agent_type", C.c_uint8 * MAX_AGENTS),
        ("agent_tile", C.c_uint8 * MAX_AGENTS),
        ("start_row", C.c_uint8 * MAX_AGENTS),
        ("start_col", C.c_uint8 * MAX_AGENTS),
        ("obs_reverse", C.c_uint8 * MAX_AGENTS),
        ("meta_hp_src", C.c_uint8 * MAX_AGENTS),
        ("n_opponents", C.c_uint8 * 2),
        ("opponents", (C.c_uint8 * MAX_AGENTS) * 2),
        ("flag_pos", (C.c_uint8 * 2) * 2),
        ("capture_pos", (C.c_uint8 * 2) * 2),
        ("spawn_pos", (C.c_uint8 * 2) * 2),
        ("flag_tile", C.c_uint8 * 2),
        ("action_delta", ((C.c_int8 * 2) * N_ACTIONS) * N_TYPES),
        ("reversed_action", C.c_uint8 * 16),
        ("type_action_mask", C.c_uint8 * N_TYPES),
        ("chan_lut", (C.c_uint8 * 16) * 2),
        ("grid_template", C.c_uint8 * MAX_CELLS),
        ("hp_float", C.c_int32),
        ("reserved0", C.c_int32),
        ("hp_max_f", C.c_double * N_TYPES),
        ("damage_f", C.c_double * N_TYPES),
        ("damage_boosted_f", C.c_double * N_TYPES),
        ("heal_f", C.c_double),
        ("vault_cost_f", C.c_double),
        ("vault_min_f", C.c_double),
    ]


# gridworld_ctf.py:100-145 — rows: U, D, R, L, no-op, then the four "second" actions
_UNIT = [(-1, 0), (1, 0), (0, 1), (0, -1)]
ACTION_DELTAS = {
    0: _UNIT + [(0, 0)] * 5,
    1: _UNIT + [(0, 0)] * 5,
    2: _UNIT + [(0, 0)] + [(2 * r, 2 * c) for r, c in _UNIT],
    3: _UNIT + [(0, 0)] + _UNIT,
}

# gridworld_ctf.py:147-196
REVERSED_ACTION_MAP = {
    None: [1, 0, 3, 2, 4, 6, 5, 8, 7],
    0: [1, 0, 2, 3, 4, 6, 5, 7, 8],
    1: [0, 1, 3, 2, 4, 5, 6, 8, 7],
    2: [2, 3, 0, 1, 4, 7, 8, 5, 6],
}

AGENT_TYPE_ACTION_MASK = {0: 1, 1: 1, 2: 0, 3: 0}  # gridworld_ctf.py:218-223
AGENT_TYPE_TILE_MAP = {0: {0: 4, 1: 8}, 1: {0: 5, 1: 9}, 2: {0: 6, 1: 10}, 3: {0: 7, 1: 11}}  # :256-261
FLAG_TILE_MAP = {0: 12, 1: 13}  # :267-270

DEFAULT_AGENT_CONFIG = {0: {"team": 0, "type": 0}, 1: {"team": 1, "type": 0}}
DEFAULT_AGENT_TYPE_HP = {0: 8, 1: 6, 2: 4, 3: 4}
DEFAULT_AGENT_TYPE_DAMAGE = {0: 1, 1: 0.5, 2: 1, 3: 1}


class CompiledEnv:
    """Result of :func:`compile_config`: the POD plus the Python-side attributes callers read."""

    def __init__(self):
        self.cfg = CtfConfig()
        self.kwargs = {}


def _hp_scale(values) -> int:
    """Smallest power-of-two scale making every HP quantity an integer."""
    for shift in range(0, 11):
        s = 1 << shift
        if all((Fraction(v) * s).denominator == 1 for v in values):
            return s
    raise ValueError(
        "HP/damage/heal/vault quantities must be dyadic rationals with denominator <= 1024 "
        f"for the fixed-point device state; got {list(values)}"
    )


def load_scenario_grid(scenario: dict, agent_tiles: dict, n_agents: int) -> np.ndarray:
    """gridworld_ctf.py:352-381 — the initial grid (numpy slice clipping applies)."""
    g = scenario["GRID_SIZE"]
    grid = np.zeros((g, g), dtype=np.uint8)
    for slc in scenario["BLOCK_TILE_SLICES"]:
        grid[slc] = 1
    for slc in scenario["DESTRUCTIBLE_TILE_SLICES"]:
        grid[slc] = 2
    grid[scenario["FLAG_POSITIONS"][0]] = FLAG_TILE_MAP[0]
    grid[scenario["FLAG_POSITIONS"][1]] = FLAG_TILE_MAP[1]
    for i in range(n_agents):
        grid[scenario["AGENT_STARTING_POSITIONS"][i]] = agent_tiles[i]
    return grid


def tiles_used(grid: np.ndarray, agent_types: dict) -> list[int]:
    """gridworld_ctf.py:488-499, including CPython's ``list(set(...))`` ordering."""
    tiles = [int(x) for x in np.unique(grid) if x != 0]
    if 2 in tiles and 3 in agent_types.values():
        tiles.extend([3])
    tiles += [8 + t for t in agent_types.values()]
    return list(set(tiles))


def compile_config(
    AGENT_CONFIG=None,
    SCENARIO=None,
    GAME_STEPS=256,
    GRID_SIZE=10,
    ENABLE_OBSTACLES=False,
    DROP_FLAG_WHEN_NO_HP=False,
    HOME_FLAG_CAPTURE=False,
    USE_EASY_CAPTURE=True,
    USE_ADJUSTED_REWARDS=False,
    MAX_BLOCK_TILE_PCT=0.2,
    LOG_METRICS=True,
    MAP_SYMMETRY_CHECK=True,
    AGENT_TYPE_HP=None,
    AGENT_HP_HEALING_PER_STEP=0.25,
    AGENT_TYPE_DAMAGE=None,
    TAG_PROBABILITY=0.75,
    GUARDIAN_DAMAGE_MULTIPLIER=5.0,
    VAULT_HP_COST=0.5,
    VAULT_MIN_HP=2.5,
    reverse_team1_actions=False,
) -> CompiledEnv:
    """Same keyword arguments and defaults as ``GridworldCtf.__init__`` (gridworld_ctf.py:19-52)."""
    AGENT_CONFIG = DEFAULT_AGENT_CONFIG if AGENT_CONFIG is None else AGENT_CONFIG
    AGENT_TYPE_HP = DEFAULT_AGENT_TYPE_HP if AGENT_TYPE_HP is None else AGENT_TYPE_HP
    AGENT_TYPE_DAMAGE = DEFAULT_AGENT_TYPE_DAMAGE if AGENT_TYPE_DAMAGE is None else AGENT_TYPE_DAMAGE
    if SCENARIO is None:
        # gridworld_ctf.py:501-567 generate_map() reads self.FLAG_POSITIONS before it exists and
        # raises AttributeError in the reference; every experiment passes a scenario.
        raise ValueError("SCENARIO is required (the reference's random map generator is broken)")

    n = len(AGENT_CONFIG)
    if not (1 <= n <= MAX_AGENTS):
        raise ValueError(f"N_AGENTS must be in 1..{MAX_AGENTS}, got {n}")
    if list(AGENT_CONFIG.keys()) != list(range(n)):
        # the reference derives OPPONENTS, TILES_USED and the metadata HP slots from the dict's insertion order
        # (gridworld_ctf.py:203-206, 393-394, 1040): only the ordered form 0..N-1 used by every experiment is supported
        raise ValueError("AGENT_CONFIG keys must be 0..N-1 in that order")
    g = int(SCENARIO["GRID_SIZE"])  # load_scenario overrides the ctor's GRID_SIZE (:359)
    if not (2 <= g <= MAX_GRID):
        raise ValueError(f"GRID_SIZE must be in 2..{MAX_GRID}, got {g}")

    out = CompiledEnv()
    cfg = out.cfg
    teams = {k: int(AGENT_CONFIG[k]["team"]) for k in AGENT_CONFIG.keys()}
    types = {k: int(AGENT_CONFIG[k]["type"]) for k in AGENT_CONFIG.keys()}
    for k in range(n):
        if teams[k] not in (0, 1) or types[k] not in (0, 1, 2, 3):
            raise ValueError(f"agent {k}: team must be 0/1 and type 0..3")
    tiles = {k: AGENT_TYPE_TILE_MAP[types[k]][teams[k]] for k in range(n)}

    # the metadata HP quirk (:1040-1041) reads agent_hp[type value]: KeyError when type >= N
    for k in range(n):
        if types[k] >= n:
            raise KeyError(
                f"agent type {types[k]} used as an agent id in get_env_metadata (gridworld_ctf.py:1041) "
                f"but N_AGENTS = {n}"
            )

    for t in (0, 1):
        x, y = SCENARIO["SPAWN_POSITIONS"][t]
        if x < 1 or y < 1:
            raise ValueError("SPAWN_POSITIONS must have row, col >= 1 (gridworld_ctf.py:773 warning)")

    grid = load_scenario_grid(SCENARIO, tiles, n)
    used = tiles_used(grid, types)
    if len(used) + 1 > 14:
        raise ValueError("more than 14 observation channels")

    # ---- scalars
    cfg.grid_size = g
    cfg.n_agents = n
    cfg.n_channels = len(used) + 1
    cfg.game_steps = int(GAME_STEPS)
    flip = SCENARIO["FLIP_AXIS"]
    if flip not in (None, 0, 1, 2):
        raise ValueError(f"FLIP_AXIS must be None, 0, 1 or 2; got {flip!r}")
    cfg.flip_axis = -1 if flip is None else int(flip)
    cfg.use_adjusted_rewards = int(bool(USE_ADJUSTED_REWARDS))
    cfg.home_flag_capture = int(bool(HOME_FLAG_CAPTURE))
    cfg.drop_flag_when_no_hp = int(bool(DROP_FLAG_WHEN_NO_HP))
    cfg.reverse_team1_actions = int(bool(reverse_team1_actions))

    # ---- rewards: constants fixed in the reference ctor (:75-80)
    reward_capture, opp_punish = 1, 0.5
    cfg.reward_step = 0.0
    cfg.reward_capture = float(reward_capture)
    cfg.reward_tag = 0.0
    cfg.capture_punish = 1.0 * reward_capture * opp_punish
    cfg.win_margin_scalar = 0.1
    cfg.loss_margin_scalar = 0.00

    # ---- HP fixed point
    for t in range(N_TYPES):
        if t not in AGENT_TYPE_HP or t not in AGENT_TYPE_DAMAGE:
            raise KeyError(f"AGENT_TYPE_HP / AGENT_TYPE_DAMAGE need an entry for type {t}")
    boosted = {t: AGENT_TYPE_DAMAGE[t] * GUARDIAN_DAMAGE_MULTIPLIER for t in range(N_TYPES)}
    quantities = (
        [AGENT_TYPE_HP[t] for t in range(N_TYPES)]
        + [AGENT_TYPE_DAMAGE[t] for t in range(N_TYPES)]
        + [boosted[t] for t in range(N_TYPES)]
        + [AGENT_HP_HEALING_PER_STEP, VAULT_HP_COST, VAULT_MIN_HP]
    )
    try:
        s = _hp_scale(quantities)
        if max(int(Fraction(v) * s) for v in quantities) > 32000:
            raise ValueError("HP quantities out of the int16 fixed-point range")
    except ValueError:
        s = 0
    if min(AGENT_TYPE_HP[t] for t in range(N_TYPES)) <= 0:
        raise ValueError("AGENT_TYPE_HP must be positive")
    cfg.hp_scale = s
    # the float fields are always filled: what the reference computes with (Python floats; int * float is exact for these)
    cfg.heal_f, cfg.vault_cost_f, cfg.vault_min_f = float(AGENT_HP_HEALING_PER_STEP), float(VAULT_HP_COST), float(VAULT_MIN_HP)
    for t in range(N_TYPES):
        cfg.hp_max_f[t] = float(AGENT_TYPE_HP[t])
        cfg.damage_f[t] = float(AGENT_TYPE_DAMAGE[t])
        cfg.damage_boosted_f[t] = float(AGENT_TYPE_DAMAGE[t] * GUARDIAN_DAMAGE_MULTIPLIER)   # :818, one multiplication
    if s:
        # every quantity is a dyadic rational: exact fixed point on the device (all shipped configurations)
        q = lambda v: int(Fraction(v) * s)  # noqa: E731
        cfg.hp_float = 0
        cfg.heal_q = q(AGENT_HP_HEALING_PER_STEP)
        cfg.vault_cost_q = q(VAULT_HP_COST)
        cfg.vault_min_q = q(VAULT_MIN_HP)
        for t in range(N_TYPES):
            cfg.hp_max_q[t] = q(AGENT_TYPE_HP[t])
            cfg.damage_q[t] = q(AGENT_TYPE_DAMAGE[t])
            cfg.damage_boosted_q[t] = q(boosted[t])
    else:
        # e.g. AGENT_HP_HEALING_PER_STEP=0.1: HP as IEEE doubles, the reference's float operations one by one
        cfg.hp_float = 1

    # u = w / 2**32 < p  <=>  w < ceil(p * 2**32)
    thr = Fraction(TAG_PROBABILITY) * (1 << 32)
    thr_i = -((-thr.numerator) // thr.denominator)
    cfg.tag_threshold = max(0, min(1 << 32, thr_i))

    cfg.zone_distance = 3       # DEFENSIVE_ZONE_DISTANCE (:87)
    cfg.guardian_distance = 3   # GUARDIAN_DEFENSE_DISTANCE (:226)
    cfg.tagging_range = 1       # GUARDIAN_TAGGING_RANGE (:227)
    cfg.max_agent_blocks = 1000  # MAX_AGENT_BLOCKS (:241)
    cfg.block_pickup_value = 1  # BLOCK_PICKUP_VALUE (:85)

    # ---- agents
    starts = SCENARIO["AGENT_STARTING_POSITIONS"]
    for k in range(n):
        cfg.agent_team[k] = teams[k]
        cfg.agent_type[k] = types[k]
        cfg.agent_tile[k] = tiles[k]
        cfg.start_row[k], cfg.start_col[k] = starts[k]
        cfg.obs_reverse[k] = int(teams[k] != 0)  # utils.py:535, ppo.py:69/87
        cfg.meta_hp_src[k] = types[k]
    half = n // 2
    opponents = {
        0: [k for k in range(n) if teams[k] == 1][:half],
        1: [k for k in range(n) if teams[k] == 0][:half],
    }
    for t in (0, 1):
        cfg.n_opponents[t] = len(opponents[t])
        if len(opponents[t]) > 4:
            raise ValueError("at most 4 opponents per team")
        for j, k in enumerate(opponents[t]):
            cfg.opponents[t][j] = k
        cfg.flag_pos[t][0], cfg.flag_pos[t][1] = SCENARIO["FLAG_POSITIONS"][t]
        cfg.capture_pos[t][0], cfg.capture_pos[t][1] = SCENARIO["CAPTURE_POSITIONS"][t]
        cfg.spawn_pos[t][0], cfg.spawn_pos[t][1] = SCENARIO["SPAWN_POSITIONS"][t]
        cfg.flag_tile[t] = FLAG_TILE_MAP[t]

    for t in range(N_TYPES):
        cfg.type_action_mask[t] = AGENT_TYPE_ACTION_MASK[t]
        for a in range(N_ACTIONS):
            cfg.action_delta[t][a][0], cfg.action_delta[t][a][1] = ACTION_DELTAS[t][a]
    for a in range(N_ACTIONS):
        cfg.reversed_action[a] = REVERSED_ACTION_MAP[flip][a]

    # ---- observation channel LUTs (standardise_state, :987-1001)
    # team-0 observers see tile codes as stored; team-1 observers see 4..7 <-> 8..11 and 12 <-> 13.
    def view_tile(observer_team: int, tile: int) -> int:
        if observer_team == 0:
            return tile
        if 4 <= tile <= 7:
            return tile + 4
        if 8 <= tile <= 11:
            return tile - 4
        if tile == 12:
            return 13
        if tile == 13:
            return 12
        return tile

    chan_of = {tile: i + 1 for i, tile in enumerate(used)}
    for team in (0, 1):
        for tile in range(16):
            cfg.chan_lut[team][tile] = chan_of.get(view_tile(team, tile), 0) if tile < 14 else 0

    flat = grid.reshape(-1)
    for i in range(g * g):
        cfg.grid_template[i] = int(flat[i])

    # ---- Python-side attributes (what ppo.py / utils.py / league_training.py read)
    out.N_AGENTS = n
    out.GRID_SIZE = g
    out.GAME_STEPS = int(GAME_STEPS)
    out.FLIP_AXIS = flip
    out.AGENT_CONFIG = AGENT_CONFIG
    out.SCENARIO = SCENARIO
    out.AGENT_TEAMS = teams
    out.AGENT_TYPES = types
    out.AGENT_TILE_MAP = tiles
    out.AGENT_TYPE_ACTION_MASK = dict(AGENT_TYPE_ACTION_MASK)
    out.AGENT_TYPE_HP = AGENT_TYPE_HP
    out.AGENT_TYPE_DAMAGE = AGENT_TYPE_DAMAGE
    out.OPPONENTS = opponents
    out.TILES_USED = used
    out.FLAG_POSITIONS = SCENARIO["FLAG_POSITIONS"]
    out.CAPTURE_POSITIONS = SCENARIO["CAPTURE_POSITIONS"]
    out.SPAWN_POSITIONS = SCENARIO["SPAWN_POSITIONS"]
    out.AGENT_STARTING_POSITIONS = SCENARIO["AGENT_STARTING_POSITIONS"]
    out.SCENARIO_NAME = SCENARIO.get("SCENARIO_NAME", "")
    out.MAP_SYMMETRY_CHECK = bool(MAP_SYMMETRY_CHECK)
    out.USE_ADJUSTED_REWARDS = bool(USE_ADJUSTED_REWARDS)
    out.REVERSED_ACTION_MAP = {k: dict(enumerate(v)) for k, v in REVERSED_ACTION_MAP.items()}
    out.initial_grid = grid
    out.n_channels = len(used) + 1
    out.meta_size = 6 + 2 * n
    return out


def env_dims(ce: CompiledEnv):
    """gridworld_ctf.py:1011-1025 get_env_dims (ACTION_SPACE = 8 at :71)."""
    c, g, n = ce.n_channels, ce.GRID_SIZE, ce.N_AGENTS
    return (c, g, g), (c - 1, g, g), (n * 2 + 6,), (n * 6 + n * 8 + 3,)
