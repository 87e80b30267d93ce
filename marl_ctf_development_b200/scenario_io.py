"""JSON round-trip of ``env_config`` dicts (scenarios.py map dict + ctor kwargs).

The reference keeps its maps as Python dicts with numpy index tuples / slices
(scenarios.py:5-624) and its experiment settings as ``TrainingConfig().env_config``
(e.g. 8_arena.py:33-63).  Those are *inputs* of the step path.  The reference
tree is not present on the GPU box, so the nine experiment configs are exported
once (tests/golden/make_golden.py, run where /root/reference exists) to
``data/experiments.json`` and rebuilt here into exactly the dict shape
``GridworldCtf(**env_config)`` takes.
"""
from __future__ import annotations

import json
import os

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "experiments.json")

_INT_KEYED = ("AGENT_CONFIG", "AGENT_TYPE_HP", "AGENT_TYPE_DAMAGE")
_SCN_INT_KEYED = ("FLAG_POSITIONS", "CAPTURE_POSITIONS", "SPAWN_POSITIONS", "AGENT_STARTING_POSITIONS")
_SCN_SLICES = ("BLOCK_TILE_SLICES", "DESTRUCTIBLE_TILE_SLICES")


def _enc_index(ix):
    if isinstance(ix, slice):
        return {"slice": [ix.start, ix.stop, ix.step]}
    return int(ix)


def _dec_index(ix):
    if isinstance(ix, dict):
        return slice(*ix["slice"])
    return int(ix)


def dump_env_config(env_config: dict) -> dict:
    out = {}
    for k, v in env_config.items():
        if k == "SCENARIO":
            scn = {}
            for sk, sv in v.items():
                if sk in _SCN_INT_KEYED:
                    scn[sk] = {str(i): [int(p[0]), int(p[1])] for i, p in sv.items()}
                elif sk in _SCN_SLICES:
                    scn[sk] = [[_enc_index(ix) for ix in entry] for entry in sv]
                else:
                    scn[sk] = sv
            out[k] = scn
        elif k in _INT_KEYED:
            out[k] = {str(i): vv for i, vv in v.items()}
        else:
            out[k] = v
    return out


def load_env_config(obj: dict) -> dict:
    out = {}
    for k, v in obj.items():
        if k == "SCENARIO":
            scn = {}
            for sk, sv in v.items():
                if sk in _SCN_INT_KEYED:
                    scn[sk] = {int(i): (int(p[0]), int(p[1])) for i, p in sv.items()}
                elif sk in _SCN_SLICES:
                    scn[sk] = [tuple(_dec_index(ix) for ix in entry) for entry in sv]
                else:
                    scn[sk] = sv
            out[k] = scn
        elif k in _INT_KEYED:
            out[k] = {int(i): vv for i, vv in v.items()}
        else:
            out[k] = v
    return out


_cache = None


def experiment_names(alternative: bool = False) -> list[str]:
    """The nine experiment scripts of the reference; with alternative=True the alt_exp/*.py variants ('alt_exp/<name>')."""
    return [k for k in _load().keys() if k.startswith("alt_exp/") == bool(alternative)]


def _load() -> dict:
    global _cache
    if _cache is None:
        with open(_DATA) as f:
            _cache = json.load(f)
    return _cache


def experiment_env_config(name: str) -> dict:
    """``env_config`` of one of the reference's experiment scripts, e.g. ``'8_arena'``."""
    data = _load()
    if name not in data:
        raise KeyError(f"unknown experiment {name!r}; have {list(data)}")
    return load_env_config(data[name])
