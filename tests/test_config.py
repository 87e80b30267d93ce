"""Config compiler against the reference's recorded known answers (no reference import needed)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, compiled
from marl_ctf_development_b200.config import compile_config, env_dims
from marl_ctf_development_b200.scenario_io import (
    dump_env_config,
    experiment_env_config,
    experiment_names,
    load_env_config,
)

# SURVEY.md §8 size table ([PROBED] on the reference)
EXPECTED = {
    "0_the_split": dict(G=11, N=4, C=8, tiles=[1, 4, 5, 8, 9, 12, 13], flip=2, types=[1, 0, 0, 0]),
    "7_gridlocked": dict(G=13, N=6, C=13, tiles=[1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13], flip=-1, types=[0, 0, 3, 2, 1, 1]),
    "8_arena": dict(G=15, N=8, C=14, tiles=list(range(1, 14)), flip=-1, types=[1, 1, 2, 2, 3, 3, 0, 0]),
}


def test_nine_experiments_present():
    assert len(experiment_names()) == 9
    alt = experiment_names(alternative=True)        # alt_exp/*.py: arena, arena_ii, jailbreak_ii, ... maps
    assert len(alt) == 8 and all(compiled(name).N_AGENTS >= 2 for name in alt)


@pytest.mark.parametrize("exp", sorted(EXPECTED))
def test_sizes_match_survey_table(exp):
    ce = compiled(exp)
    want = EXPECTED[exp]
    assert ce.GRID_SIZE == want["G"] and ce.N_AGENTS == want["N"] and ce.n_channels == want["C"]
    assert [int(t) for t in ce.TILES_USED] == want["tiles"]
    assert ce.cfg.flip_axis == want["flip"]
    assert [ce.AGENT_TYPES[i] for i in range(ce.N_AGENTS)] == want["types"]
    assert [ce.AGENT_TEAMS[i] for i in range(ce.N_AGENTS)] == [i % 2 for i in range(ce.N_AGENTS)]
    # HP x4 / damage x4 / vault columns of the table
    assert ce.cfg.hp_scale == 4
    assert list(ce.cfg.hp_max_q) == [40, 32, 32, 28]
    assert list(ce.cfg.damage_q) == [4, 2, 2, 4]
    assert (ce.cfg.vault_cost_q, ce.cfg.vault_min_q, ce.cfg.heal_q) == (5, 10, 1)
    assert ce.cfg.tag_threshold == 0xC0000000
    assert ce.cfg.game_steps == 500 and ce.cfg.use_adjusted_rewards == 1


def test_env_dims_formula():
    # env_testing.ipynb:249 records ((8, 11, 11), (7, 11, 11), (14,), (59,)) for a 4-agent arrow env
    ec = experiment_env_config("0_the_split")
    assert env_dims(compile_config(**ec)) == ((8, 11, 11), (7, 11, 11), (14,), (59,))


def test_notebook_initial_grid_arrow():
    """env_testing.ipynb:59-88 — env.grid of scn.arrow with types [0,1,0,0] right after reset."""
    ec = experiment_env_config("0_the_split")
    ec["AGENT_CONFIG"] = {0: {"team": 0, "type": 0}, 1: {"team": 1, "type": 1}, 2: {"team": 0, "type": 0}, 3: {"team": 1, "type": 0}}
    ec.update(GAME_STEPS=256, USE_ADJUSTED_REWARDS=False, HOME_FLAG_CAPTURE=True, MAP_SYMMETRY_CHECK=False)
    want = np.zeros((11, 11), dtype=np.uint8)
    want[1, 1] = 12
    want[5, 1] = want[5, 2] = 4
    want[5, 5] = want[6, 4] = want[7, 3] = want[8, 2] = want[9, 1] = want[10, 0] = 1
    want[8, 5] = 9
    want[9, 5] = 8
    want[9, 9] = 13
    assert np.array_equal(compile_config(**ec).initial_grid, want)


def test_json_round_trip_is_identity():
    for name in experiment_names():
        ec = experiment_env_config(name)
        again = load_env_config(json.loads(json.dumps(dump_env_config(ec))))
        assert repr(again) == repr(ec)


def test_reset_state_matches_reference_json_traces():
    """Reset-state content of the reference's json/*.json (utils.py:745-754), extracted by make_golden.py."""
    with open(os.path.join(GOLDEN, "json_reset_states.json")) as f:
        states = json.load(f)
    assert len(states) == 27
    for base, s in states.items():
        ce = compiled(s["experiment"])
        g = ce.initial_grid
        assert ce.GRID_SIZE == s["grid_size"], base
        assert sorted(map(list, zip(*np.where(g == 1)))) == s["block_tiles"], base
        destr = sorted([int(r), int(c), 0] for r, c in zip(*np.where(g == 2)))
        assert destr == s["destructible_tiles"], base
        agents = [[ce.AGENT_TEAMS[i], ce.AGENT_TYPES[i], *ce.AGENT_STARTING_POSITIONS[i]] for i in range(ce.N_AGENTS)]
        assert agents == s["agents"], base
        for t in (0, 1):
            assert list(ce.FLAG_POSITIONS[t]) == s["flag_pos"][str(t)], base
            assert list(ce.SPAWN_POSITIONS[t]) == s["spawn_pos"][str(t)], base


def test_rejects_what_the_reference_cannot_run():
    ec = experiment_env_config("8_arena")
    with pytest.raises(ValueError):
        compile_config(**{**ec, "SCENARIO": None})  # generate_map() raises AttributeError in the reference
    two = {0: {"team": 0, "type": 3}, 1: {"team": 1, "type": 0}}
    with pytest.raises(KeyError):
        compile_config(**{**ec, "AGENT_CONFIG": two})  # get_env_metadata reads agent_hp[3] (gridworld_ctf.py:1041)
    with pytest.raises(ValueError):
        compile_config(**{**ec, "AGENT_TYPE_HP": {0: 8, 1: 0, 2: 4, 3: 4}})  # an agent type without HP
    shuffled = {1: {"team": 1, "type": 0}, 0: {"team": 0, "type": 0}}
    with pytest.raises(ValueError):
        compile_config(**{**ec, "AGENT_CONFIG": shuffled})  # insertion order matters in the reference


def test_non_dyadic_hp_quantities_switch_to_float_hp():
    """Quantities that are not dyadic rationals (heal 0.1) cannot live in exact fixed point: HP becomes IEEE doubles and the
    device performs the reference's float operations one by one (cfg.hp_float); every shipped config stays fixed point."""
    ec = experiment_env_config("8_arena")
    assert compile_config(**ec).cfg.hp_float == 0 and compile_config(**ec).cfg.hp_scale == 4
    ce = compile_config(**{**ec, "AGENT_HP_HEALING_PER_STEP": 0.1})
    assert ce.cfg.hp_float == 1 and ce.cfg.hp_scale == 0 and ce.cfg.heal_f == 0.1
    assert list(ce.cfg.hp_max_f) == [10.0, 8.0, 8.0, 7.0]
    ce = compile_config(**{**ec, "AGENT_TYPE_DAMAGE": {0: 0.3, 1: 0.7, 2: 1, 3: 1}, "GUARDIAN_DAMAGE_MULTIPLIER": 2.7})
    assert ce.cfg.hp_float == 1 and ce.cfg.damage_boosted_f[1] == 0.7 * 2.7    # one float multiplication, like :818


def test_default_hp_configs_compile():
    # experiments 1-5 use the ctor-default HP {8,6,4,4}, damage {1,.5,1,1}, vault cost 0.5
    ce = compiled("1_fence")
    assert list(ce.cfg.hp_max_q) == [32, 24, 16, 16] and list(ce.cfg.damage_q) == [4, 2, 4, 4]
    assert list(ce.cfg.damage_boosted_q) == [20, 10, 20, 20] and ce.cfg.vault_cost_q == 2
