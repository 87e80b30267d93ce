"""Constructor-keyword variations of the reference ctor (gridworld_ctf.py:19-52) beyond the nine experiment scripts.

Each case is (name, base experiment, overrides, policy).  Used by the oracle-vs-reference test (CPU, where the
reference exists) and by the CUDA-vs-oracle test (GPU box), so every keyword the step path reads is pinned on
both sides.
"""

def _agents(types, teams=None):
    teams = teams if teams is not None else [i % 2 for i in range(len(types))]
    return {i: {"team": teams[i], "type": types[i]} for i in range(len(types))}


KWARG_CASES = [
    ("home_flag_capture", "8_arena", {"HOME_FLAG_CAPTURE": True}, "seek"),
    ("drop_flag_when_no_hp", "8_arena", {"DROP_FLAG_WHEN_NO_HP": True}, "seek"),
    ("plain_rewards", "7_gridlocked", {"USE_ADJUSTED_REWARDS": False}, "seek"),
    ("always_tag", "0_the_split", {"TAG_PROBABILITY": 1.0}, "seek"),
    ("never_tag", "0_the_split", {"TAG_PROBABILITY": 0.0}, "seek"),
    ("odd_tag_probability", "8_arena", {"TAG_PROBABILITY": 0.3}, "seek"),
    ("short_game", "5_skittles", {"GAME_STEPS": 37}, "builder"),
    ("ctor_default_hp", "8_arena", {"AGENT_TYPE_HP": {0: 8, 1: 6, 2: 4, 3: 4}, "AGENT_TYPE_DAMAGE": {0: 1, 1: 0.5, 2: 1, 3: 1}, "VAULT_HP_COST": 0.5}, "builder"),
    ("fine_hp_grid", "7_gridlocked", {"AGENT_TYPE_HP": {0: 5.125, 1: 6.5, 2: 9.375, 3: 4}, "AGENT_TYPE_DAMAGE": {0: 0.625, 1: 0.375, 2: 1.5, 3: 0.875},
                                      "AGENT_HP_HEALING_PER_STEP": 0.125, "VAULT_HP_COST": 0.875, "VAULT_MIN_HP": 1.25, "GUARDIAN_DAMAGE_MULTIPLIER": 2.5}, "builder"),
    ("no_damage_type", "8_arena", {"AGENT_TYPE_DAMAGE": {0: 1, 1: 0.5, 2: 0, 3: 1}}, "seek"),
    ("big_heal_small_vault_gate", "7_gridlocked", {"AGENT_HP_HEALING_PER_STEP": 2.0, "VAULT_MIN_HP": 0.0, "VAULT_HP_COST": 3.0}, "builder"),
    ("six_agents_on_arena", "8_arena", {"AGENT_CONFIG": _agents([3, 2, 1, 0, 0, 3]), "MAP_SYMMETRY_CHECK": False}, "builder"),
    ("five_agents_truncated_opponents", "8_arena", {"AGENT_CONFIG": _agents([0, 1, 2, 3, 0]), "MAP_SYMMETRY_CHECK": False}, "seek"),
    ("uneven_teams", "8_arena", {"AGENT_CONFIG": _agents([0, 1, 2, 3, 1, 2], teams=[0, 0, 0, 0, 1, 1]), "MAP_SYMMETRY_CHECK": False}, "seek"),
    ("four_scouts_tiny_tile_set", "0_the_split", {"AGENT_CONFIG": _agents([0, 0, 0, 0]), "MAP_SYMMETRY_CHECK": False}, "seek"),
    ("all_miners", "2_jailbreak", {"AGENT_CONFIG": _agents([3, 3, 3, 3]), "MAP_SYMMETRY_CHECK": False}, "builder"),
    # every hit is lethal and always lands: respawn storms, several lethal hits in one actor's turn, carriers killed
    ("one_hit_kills_arena", "8_arena", {"TAG_PROBABILITY": 1.0, "AGENT_TYPE_HP": {0: 1, 1: 1, 2: 1, 3: 1},
                                        "AGENT_TYPE_DAMAGE": {0: 1, 1: 1, 2: 1, 3: 1}, "VAULT_MIN_HP": 0.0, "VAULT_HP_COST": 0.25}, "seek"),
    ("one_hit_kills_split", "0_the_split", {"TAG_PROBABILITY": 1.0, "AGENT_TYPE_HP": {0: 0.5, 1: 0.5, 2: 0.5, 3: 0.5},
                                            "AGENT_TYPE_DAMAGE": {0: 0.5, 1: 0.25, 2: 0.5, 3: 0.5}, "GUARDIAN_DAMAGE_MULTIPLIER": 2.0}, "seek"),
    ("glass_cannons_gridlocked", "7_gridlocked", {"TAG_PROBABILITY": 0.9, "AGENT_TYPE_HP": {0: 1.5, 1: 1, 2: 2, 3: 1},
                                                  "AGENT_TYPE_DAMAGE": {0: 1, 1: 0.5, 2: 1, 3: 2}, "AGENT_HP_HEALING_PER_STEP": 0.5}, "builder"),
    # HP quantities that are not dyadic rationals: the backend switches to IEEE-double HP (cfg.hp_float) and has to follow the
    # reference's Python float arithmetic operation by operation (0.1 + 0.1 + 0.1 != 0.3)
    ("float_hp_heal_tenth", "8_arena", {"AGENT_HP_HEALING_PER_STEP": 0.1}, "seek"),
    ("float_hp_everything", "8_arena", {"AGENT_HP_HEALING_PER_STEP": 0.1, "AGENT_TYPE_DAMAGE": {0: 0.3, 1: 0.7, 2: 0.9, 3: 0.35},
                                        "VAULT_HP_COST": 0.45, "VAULT_MIN_HP": 1.1, "AGENT_TYPE_HP": {0: 3.3, 1: 2.9, 2: 4.1, 3: 2.2},
                                        "GUARDIAN_DAMAGE_MULTIPLIER": 2.7}, "seek"),
    ("float_hp_gridlocked_builder", "7_gridlocked", {"AGENT_TYPE_DAMAGE": {0: 0.3, 1: 0.2, 2: 0.6, 3: 0.7}, "VAULT_HP_COST": 0.3,
                                                     "VAULT_MIN_HP": 0.7, "AGENT_HP_HEALING_PER_STEP": 0.05}, "builder"),
    ("float_hp_one_hit_kills", "0_the_split", {"TAG_PROBABILITY": 1.0, "AGENT_TYPE_HP": {0: 0.3, 1: 0.3, 2: 0.3, 3: 0.3},
                                               "AGENT_TYPE_DAMAGE": {0: 0.3, 1: 0.1, 2: 0.3, 3: 0.3}, "GUARDIAN_DAMAGE_MULTIPLIER": 3.0}, "seek"),
]

CASE_IDS = [c[0] for c in KWARG_CASES]


def build_env_config(base_config: dict, overrides: dict) -> dict:
    ec = dict(base_config)
    ec.update(overrides)
    return ec
