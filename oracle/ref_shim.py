"""TEST INFRASTRUCTURE — import shim for the unmodified Python reference.

Loads ``/root/reference/gridworld_ctf.py`` (and the experiment scripts) in this
container, with empty stand-ins for the plotting / Ray modules it imports but
does not need on the step path, and injects the counter-based draws of
``marl_ctf_development_b200.draws`` at the reference's three RNG sites
(gridworld_ctf.py:740 ``random.shuffle``, :815 ``np.random.rand``, :771
``np.random.randint``) so that its results can be compared bit for bit with
the oracle and the CUDA path.

The reference does not exist on the GPU box: only tests that are skipped when
``/root/reference`` is absent, and ``tests/golden/make_golden.py``, use this.
Nothing under ``marl_ctf_development_b200/`` imports it.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import runpy
import sys
import types

import numpy as np

SRC_ROOT = os.environ.get("CTF_REFERENCE_ROOT", "/root/reference")
# oracle/build_ref.py byte-compiles the reference's own files into oracle/_ref/*.refc (git-ignored build output, no
# sources): the unmodified reference then also runs where /root/reference does not exist (the GPU box)
PYC_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _source_available() -> bool:
    return os.path.isfile(os.path.join(SRC_ROOT, "gridworld_ctf.py"))


def compiled_available() -> bool:
    return os.path.isfile(os.path.join(PYC_ROOT, "gridworld_ctf.refc"))


REF_ROOT = SRC_ROOT if _source_available() or not compiled_available() else PYC_ROOT
REF_SUFFIX = ".py" if REF_ROOT == SRC_ROOT else ".refc"


class _CompiledRefFinder(importlib.abc.MetaPathFinder):
    """Imports ``import gridworld_ctf`` etc. from oracle/_ref/<name>.refc (byte code written by py_compile)."""

    def find_spec(self, fullname, path=None, target=None):
        cand = os.path.join(PYC_ROOT, fullname + ".refc")
        if "." in fullname or not os.path.isfile(cand):
            return None
        loader = importlib.machinery.SourcelessFileLoader(fullname, cand)
        return importlib.util.spec_from_file_location(fullname, cand, loader=loader)

EXPERIMENTS = (
    "0_the_split",
    "1_fence",
    "2_jailbreak",
    "3_one_way_out",
    "4_keyhole",
    "5_skittles",
    "6_the_wall",
    "7_gridlocked",
    "8_arena",
)


def available() -> bool:
    """The reference can be imported here: from its source tree, or from the byte-compiled oracle/_ref."""
    return os.path.isfile(os.path.join(REF_ROOT, "gridworld_ctf" + REF_SUFFIX))


def source_tree_available() -> bool:
    """The reference's source tree itself is here (needed by the fixture generators that read json/*.json)."""
    return REF_ROOT == SRC_ROOT and _source_available()


class _Anything(types.ModuleType):
    """A module whose every attribute is a no-op callable / identity decorator."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _noop(*args, **kwargs):
            if len(args) == 1 and callable(args[0]) and not kwargs:
                return args[0]  # used as a decorator (ray.remote)
            return None

        return _noop


_STUB_ROOTS = ("IPython", "matplotlib", "seaborn", "imageio", "wandb", "distutils")


class _EagerRemote:
    """``@ray.remote`` stand-in: ``f.remote(*a, **k)`` runs f at once and returns its result."""

    def __init__(self, fn):
        self._fn = fn

    def remote(self, *args, **kwargs):
        return self._fn(*args, **kwargs)

    def __call__(self, *args, **kwargs):
        return self._fn(*args, **kwargs)


def _ray_stub():
    """Ray is absent here; the reference's callers only need remote / get / put / init / shutdown
    (league_training.py:16-18, 384-386, 683-687; ppo.py:264-266, 349-359).  Tasks run eagerly, in submission order."""
    mod = types.ModuleType("ray")
    mod.remote = lambda fn: _EagerRemote(fn)
    mod.get = lambda x: x
    mod.put = lambda x: x
    mod.init = lambda *a, **k: None
    mod.shutdown = lambda *a, **k: None
    return mod


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serves an empty stand-in for any (sub)module of a package that is absent here."""

    def __init__(self, roots):
        self.roots = roots

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _Anything(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def _install_stubs():
    missing = []
    for name in _STUB_ROOTS:
        if importlib.util.find_spec(name) is None:
            missing.append(name)
    if missing:
        sys.meta_path.append(_StubFinder(tuple(missing)))
    if importlib.util.find_spec("ray") is None and "ray" not in sys.modules:
        sys.modules["ray"] = _ray_stub()


@contextlib.contextmanager
def _in_ref_dir():
    """The reference ctor opens os.getcwd() + '/img/*.png' (gridworld_ctf.py:319)."""
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


class _NoSprites:
    """Stands in for PIL.Image inside gridworld_ctf where the sprite folder is absent (oracle/_ref holds no assets):
    the ctor only stores the 24 opened images for render_image (gridworld_ctf.py:319-347, out of scope)."""

    @staticmethod
    def open(path):
        return None


_modules = {}


def reference_modules():
    """Returns (gridworld_ctf module, scenarios module) of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if not _modules:
        _install_stubs()
        if REF_SUFFIX == ".py":
            if REF_ROOT not in sys.path:
                sys.path.insert(0, REF_ROOT)
        elif not any(isinstance(f, _CompiledRefFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _CompiledRefFinder())
        with _in_ref_dir():
            _modules["gw"] = importlib.import_module("gridworld_ctf")
            _modules["scn"] = importlib.import_module("scenarios")
        if not os.path.isdir(os.path.join(REF_ROOT, "img")):
            _modules["gw"].Image = _NoSprites
    return _modules["gw"], _modules["scn"]


def caller_modules() -> dict:
    """The reference's callers of the step path, unmodified: ppo (PPOTrainer), utils (duel, duel_json),
    league_training (LeagueTrainer), agent_network (Agent), metrics_logger (MetricsLogger)."""
    reference_modules()
    if "ppo" not in _modules:
        with _in_ref_dir():
            for name in ("agent_network", "metrics_logger", "utils", "ppo", "league_training"):
                _modules[name] = importlib.import_module(name)
    return {k: _modules[k] for k in ("agent_network", "metrics_logger", "utils", "ppo", "league_training")}


def experiment_env_config(name: str) -> dict:
    """``TrainingConfig().env_config`` of an experiment script (e.g. '8_arena')."""
    reference_modules()
    with _in_ref_dir():
        ns = runpy.run_path(os.path.join(REF_ROOT, name + REF_SUFFIX), run_name="ctf_ref_shim")
    return ns["TrainingConfig"]().env_config


def alt_experiment_names() -> list:
    """Names of the alternative experiment scripts (alt_exp/*.py, "not used in final report"); source tree only."""
    d = os.path.join(SRC_ROOT, "alt_exp")
    return sorted(f[:-3] for f in os.listdir(d) if f.endswith(".py")) if source_tree_available() and os.path.isdir(d) else []


def alt_experiment_env_config(name: str) -> dict:
    reference_modules()
    with _in_ref_dir():
        ns = runpy.run_path(os.path.join(SRC_ROOT, "alt_exp", name + ".py"), run_name="ctf_ref_shim")
    return ns["TrainingConfig"]().env_config


def make_reference_env(env_config: dict):
    """Unmodified reference env, global RNGs untouched (caller seeds them)."""
    gw, _ = reference_modules()
    with _in_ref_dir():
        return gw.GridworldCtf(**env_config)


# --------------------------------------------------------------------------------------
# Draw injection
# --------------------------------------------------------------------------------------
class _InjectedRandom:
    """Stands in for the ``random`` module inside gridworld_ctf (only shuffle is used)."""

    def __init__(self, hub):
        self._hub = hub

    def shuffle(self, arr):
        self._hub.shuffle(arr)


class _InjectedNpRandom:
    def __init__(self, hub):
        self._hub = hub

    def rand(self, *shape):
        assert not shape
        return self._hub.rand()

    def randint(self, *args, **kwargs):
        assert len(args) == 1 and not kwargs
        return self._hub.randint(args[0])

    def __getattr__(self, name):
        return getattr(np.random, name)


class _NumpyProxy:
    """``np`` as seen from gridworld_ctf: numpy with ``.random`` redirected."""

    def __init__(self, hub):
        self.random = _InjectedNpRandom(hub)

    def __getattr__(self, name):
        return getattr(np, name)


class _Hub:
    """Routes the module-level RNG calls to the env that is currently stepping."""

    def __init__(self):
        self.env = None

    def shuffle(self, arr):
        self.env._inj_shuffle(arr)

    def rand(self):
        return self.env._inj_rand()

    def randint(self, k):
        return self.env._inj_randint(k)


_hub = _Hub()
_injected_cls = None


def injected_env_class():
    """Subclass of the reference ``GridworldCtf`` whose draws are the Philox site draws."""
    global _injected_cls
    if _injected_cls is not None:
        return _injected_cls
    gw, _ = reference_modules()
    from marl_ctf_development_b200 import draws

    # redirect the module-level names used at gridworld_ctf.py:740, :771, :815
    gw.random = _InjectedRandom(_hub)
    gw.np = _NumpyProxy(_hub)

    class InjectedGridworldCtf(gw.GridworldCtf):
        def __init__(self, *args, seed=0, env_id=0, reset_advances="episode", **kwargs):
            # reset_advances="episode": every reset() starts the next episode of env `env_id` (how one GPU env behaves).
            # reset_advances="env": the ctor's reset is (env_id, episode 0); the k-th later reset() becomes episode 1 of
            # env `env_id + k - 1` — a caller that plays one duel after the other on ONE env object (league_training.py
            # submits number_of_duels tasks with the same env) then visits the same draws as a GPU batch whose env
            # env_id + k - 1 plays duel k after the batch-wide reset.
            self._inj_seed = int(seed)
            self._inj_env_id = int(env_id)
            self._inj_first_env_id = int(env_id)
            self._inj_reset_advances = reset_advances
            self._inj_resets = -1
            self._inj_episode = -1
            self._inj_words = None
            self._inj_shuffles = 0
            self.draw_log = []
            with _in_ref_dir():
                super().__init__(*args, **kwargs)

        # -- draw sources -----------------------------------------------------------
        def _inj_shuffle(self, arr):
            self._inj_shuffles += 1
            if self._inj_shuffles == 1:  # move order (:861); the heal shuffle (:844) leaves arr alone
                arr[:] = draws.move_order(self._inj_words, self.N_AGENTS)

        def _inj_rand(self):
            u = draws.tag_roll_uniform(self._inj_words, self._inj_actor, self._inj_slot)
            self._inj_last = (self._inj_actor, self._inj_slot)
            self._inj_slot += 1
            return u

        def _inj_randint(self, k):
            actor, slot = self._inj_last
            return draws.respawn_pick(self._inj_words, actor, slot, int(k))

        # -- hooks ------------------------------------------------------------------
        def reset(self):
            self._inj_resets += 1
            if self._inj_reset_advances == "env" and self._inj_resets >= 1:
                self._inj_env_id = self._inj_first_env_id + self._inj_resets - 1
                self._inj_episode = 1
            else:
                self._inj_episode += 1
            _hub.env = self
            return super().reset()

        def tagging_logic(self, agent_idx):
            self._inj_actor = int(agent_idx)
            self._inj_slot = 0
            return super().tagging_logic(agent_idx)

        def step(self, actions):
            _hub.env = self
            self._inj_words = draws.step_words(
                self._inj_seed, self._inj_env_id, self._inj_episode, self.env_step_count + 1
            )
            self._inj_shuffles = 0
            return super().step(actions)

    _injected_cls = InjectedGridworldCtf
    return _injected_cls


def make_injected_env(env_config: dict, seed: int = 0, env_id: int = 0, reset_advances: str = "episode"):
    return injected_env_class()(**env_config, seed=seed, env_id=env_id, reset_advances=reset_advances)


# --------------------------------------------------------------------------------------
# State extraction in the layout the oracle / CUDA path use
# --------------------------------------------------------------------------------------
def snapshot(env, hp_scale: int) -> dict:
    """Comparable state of a reference env (numpy arrays, exact integer HP)."""
    n = env.N_AGENTS
    pos = np.array([env.agent_positions[i] for i in range(n)], dtype=np.uint8)
    hp = np.array([int(round(env.agent_hp[i] * hp_scale)) for i in range(n)], dtype=np.int32)
    for i in range(n):
        assert hp_scale == 0 or hp[i] == env.agent_hp[i] * hp_scale, "HP not exact in fixed point"
    return {
        "grid": env.grid.astype(np.uint8).copy(),
        "pos": pos,
        "hp_q": hp,
        "hp": np.array([float(env.agent_hp[i]) for i in range(n)], dtype=np.float64),   # compared bit for bit when cfg.hp_float
        "has_flag": env.has_flag.astype(np.uint8).copy(),
        "inventory": np.array([env.block_inventory[i] for i in range(n)], dtype=np.int32),
        "step": int(env.env_step_count),
        "captures": np.array(
            [env.metrics["team_flag_captures"][0], env.metrics["team_flag_captures"][1]], dtype=np.int64
        ),
    }


def observations(env) -> tuple[np.ndarray, np.ndarray]:
    """What the callers feed the policy (ppo.py:69-70, utils.py:534-543): float32 obs [N,C,G,G], meta [N,M]."""
    n = env.N_AGENTS
    obs = np.stack(
        [env.standardise_state(i, reverse_grid=(env.AGENT_TEAMS[i] != 0))[0] for i in range(n)]
    ).astype(np.float32)
    meta = np.stack([env.get_env_metadata(i)[0] for i in range(n)]).astype(np.float32)
    return obs, meta


def agent_metrics(env) -> np.ndarray:
    """Agent-level counters as int64 [13, N] in ctf_metric order (gridworld_ctf.py:456-468)."""
    from marl_ctf_development_b200.config import METRIC_NAMES

    n = env.N_AGENTS
    out = np.zeros((len(METRIC_NAMES), n), dtype=np.int64)
    for m, name in enumerate(METRIC_NAMES):
        d = env.metrics["agent_" + name]
        for i in range(n):
            out[m, i] = int(d[i]) if i in d else 0
    return out
