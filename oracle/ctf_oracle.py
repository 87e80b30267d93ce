"""TEST INFRASTRUCTURE — ctypes wrapper of the CPU oracle (oracle/ctf_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may
import this.  The product path (marl_ctf_development_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from marl_ctf_development_b200.config import N_METRICS, CompiledEnv, CtfConfig, compile_config

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libctf_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ctf_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ctf_b200.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH))
        or (os.path.exists(hdr) and os.path.getmtime(hdr) > os.path.getmtime(_LIB_PATH))
    )
    if stale:
        subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ctf_oracle_env_size.restype = C.c_size_t
        L.ctf_oracle_init.argtypes = [C.c_void_p, C.POINTER(CtfConfig), C.c_uint64, C.c_uint32]
        L.ctf_oracle_reset.argtypes = [C.c_void_p]
        L.ctf_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ctf_oracle_observe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ctf_oracle_standardise_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ctf_oracle_metadata.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ctf_oracle_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.ctf_oracle_set_state.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ctf_oracle_f64_to_f16_to_f32.argtypes = [C.c_double]
        L.ctf_oracle_f64_to_f16_to_f32.restype = C.c_float
        L.ctf_oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ctf_oracle_run_baseline.argtypes = [C.POINTER(CtfConfig), C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int]
        L.ctf_oracle_run_baseline.restype = C.c_double
        L.ctf_oracle_batch_create.argtypes = [C.POINTER(CtfConfig), C.c_int, C.c_uint64, C.c_uint32]
        L.ctf_oracle_batch_create.restype = C.c_void_p
        L.ctf_oracle_batch_destroy.argtypes = [C.c_void_p]
        L.ctf_oracle_batch_reset.argtypes = [C.c_void_p]
        L.ctf_oracle_batch_step.argtypes = [C.c_void_p] * 4
        L.ctf_oracle_batch_observe.argtypes = [C.c_void_p] * 5
        L.ctf_oracle_batch_get_state.argtypes = [C.c_void_p] * 9
        L.ctf_oracle_batch_set_state.argtypes = [C.c_void_p] * 7
        L.ctf_oracle_batch_run.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int]
        L.ctf_oracle_batch_run.restype = C.c_double
        L.ctf_oracle_observe_fast.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ctf_oracle_get_hp_f.argtypes = [C.c_void_p, C.c_void_p]
        L.ctf_oracle_batch_get_hp_f.argtypes = [C.c_void_p, C.c_void_p]
        L.ctf_oracle_batch_set_hp_f.argtypes = [C.c_void_p, C.c_void_p]
        L.ctf_oracle_take_faults.argtypes = [C.c_void_p]
        L.ctf_oracle_take_faults.restype = C.c_uint32
        L.ctf_oracle_batch_take_faults.argtypes = [C.c_void_p]
        L.ctf_oracle_batch_take_faults.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """One sequential env with the reference's semantics; ``compiled`` from config.compile_config."""

    def __init__(self, compiled: CompiledEnv | None = None, seed: int = 0, env_id: int = 0, **env_config):
        self.ce = compiled if compiled is not None else compile_config(**env_config)
        self.L = lib()
        self._buf = C.create_string_buffer(self.L.ctf_oracle_env_size())
        self._e = C.cast(self._buf, C.c_void_p)
        self.N = self.ce.N_AGENTS
        self.G = self.ce.GRID_SIZE
        self.Cn = self.ce.n_channels
        self.M = 6 + 2 * self.N
        self.L.ctf_oracle_init(self._e, C.byref(self.ce.cfg), seed, env_id)

    def reset(self):
        self.L.ctf_oracle_reset(self._e)

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.uint8))
        assert a.shape == (self.N,)
        r = np.zeros(self.N, dtype=np.float32)
        d = np.zeros(1, dtype=np.uint8)
        self.L.ctf_oracle_step(self._e, _p(a), _p(r), _p(d))
        return r, bool(d[0])

    def observe(self, reverse_flags=None, u8=False):
        obs = np.zeros((self.N, self.Cn, self.G, self.G), dtype=np.uint8 if u8 else np.float32)
        meta = np.zeros((self.N, self.M), dtype=np.float32)
        rf = None if reverse_flags is None else np.ascontiguousarray(np.asarray(reverse_flags, dtype=np.uint8))
        self.L.ctf_oracle_observe(
            self._e, None if rf is None else _p(rf), None if u8 else _p(obs), _p(obs) if u8 else None, _p(meta)
        )
        return obs, meta

    def observe_fast(self):
        """The baseline legs' observation writer (must equal observe())."""
        obs = np.full((self.N, self.Cn, self.G, self.G), 7.0, dtype=np.float32)
        meta = np.zeros((self.N, self.M), dtype=np.float32)
        self.L.ctf_oracle_observe_fast(self._e, _p(obs), _p(meta))
        return obs, meta

    def standardise_state(self, agent_idx, reverse_grid=False):
        out = np.zeros((1, self.Cn, self.G, self.G), dtype=np.uint8)
        self.L.ctf_oracle_standardise_state(self._e, int(agent_idx), int(bool(reverse_grid)), _p(out))
        return out

    def get_env_metadata(self, agent_idx):
        out = np.zeros((1, self.M), dtype=np.float32)
        self.L.ctf_oracle_metadata(self._e, int(agent_idx), _p(out))
        return out

    def state(self) -> dict:
        n, g = self.N, self.G
        grid = np.zeros((g, g), dtype=np.uint8)
        pos = np.zeros((n, 2), dtype=np.int32)
        hp = np.zeros(n, dtype=np.int32)
        flag = np.zeros(n, dtype=np.uint8)
        inv = np.zeros(n, dtype=np.int32)
        sc = np.zeros(5, dtype=np.int32)
        stats = np.zeros((N_METRICS, n), dtype=np.uint32)
        visits = np.zeros((n, g, g), dtype=np.uint8)
        self.L.ctf_oracle_get_state(self._e, _p(grid), _p(pos), _p(hp), _p(flag), _p(inv), _p(sc), _p(stats), _p(visits))
        hp_f = np.zeros(n, dtype=np.float64)
        self.L.ctf_oracle_get_hp_f(self._e, _p(hp_f))
        return {
            "grid": grid,
            "pos": pos.astype(np.uint8),
            "hp_q": hp,
            "hp": hp_f,     # agent_hp as floats (what is compared when cfg.hp_float; hp_q / hp_scale otherwise)
            "has_flag": flag,
            "inventory": inv,
            "step": int(sc[0]),
            "episode": int(sc[1]),
            "captures": sc[2:4].astype(np.int64),
            "done": bool(sc[4]),
            "stats": stats.astype(np.int64),
            "visits": visits,
        }

    def set_state(self, grid, pos, hp_q, has_flag, inventory, step, episode, captures, done=False):
        grid = np.ascontiguousarray(np.asarray(grid, dtype=np.uint8))
        pos = np.ascontiguousarray(np.asarray(pos, dtype=np.int32))
        hp = np.ascontiguousarray(np.asarray(hp_q, dtype=np.int32))
        flag = np.ascontiguousarray(np.asarray(has_flag, dtype=np.uint8))
        inv = np.ascontiguousarray(np.asarray(inventory, dtype=np.int32))
        sc = np.array([step, episode, captures[0], captures[1], int(done)], dtype=np.int32)
        self.L.ctf_oracle_set_state(self._e, _p(grid), _p(pos), _p(hp), _p(flag), _p(inv), _p(sc))


class OracleBatch:
    """B sequential oracle envs with global ids env_id_base + b (mirrors GridworldCtfGPU's batch semantics)."""

    def __init__(self, compiled: CompiledEnv, num_envs: int, seed: int = 0, env_id_base: int = 0):
        self.ce = compiled
        self.L = lib()
        self.B = int(num_envs)
        self.N, self.G, self.Cn = compiled.N_AGENTS, compiled.GRID_SIZE, compiled.n_channels
        self.M = 6 + 2 * self.N
        self._h = C.c_void_p(self.L.ctf_oracle_batch_create(C.byref(compiled.cfg), self.B, seed, env_id_base))

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.ctf_oracle_batch_destroy(self._h)
            self._h = None

    def reset(self):
        self.L.ctf_oracle_batch_reset(self._h)

    def step(self, actions):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.uint8))
        assert a.shape == (self.B, self.N)
        r = np.zeros((self.B, self.N), dtype=np.float32)
        d = np.zeros(self.B, dtype=np.uint8)
        self.L.ctf_oracle_batch_step(self._h, _p(a), _p(r), _p(d))
        return r, d

    def observe(self, reverse_flags=None, u8=False):
        obs = np.zeros((self.B, self.N, self.Cn, self.G, self.G), dtype=np.uint8 if u8 else np.float32)
        meta = np.zeros((self.B, self.N, self.M), dtype=np.float32)
        rf = None if reverse_flags is None else np.ascontiguousarray(np.asarray(reverse_flags, dtype=np.uint8))
        self.L.ctf_oracle_batch_observe(
            self._h, None if rf is None else _p(rf), None if u8 else _p(obs), _p(obs) if u8 else None, _p(meta)
        )
        return obs, meta

    def take_faults(self) -> int:
        """CTF_FAULT_* bits raised by any env since the last call (mirrors GridworldCtfGPU.take_faults)."""
        return int(self.L.ctf_oracle_batch_take_faults(self._h))

    def run(self, steps: int, seed: int, with_obs, n_threads: int) -> float:
        """Threaded continuation with uniform random actions (CPU legs of bench.py). Returns a checksum.

        with_obs: 0 = step() only; 1 / True = observations through the literal restatement of the reference's
        standardise_state; 2 = through the tuned writer (same bytes)."""
        return float(self.L.ctf_oracle_batch_run(self._h, int(steps), int(seed), int(with_obs), int(n_threads)))

    def state(self) -> dict:
        B, n, g = self.B, self.N, self.G
        grid = np.zeros((B, g, g), dtype=np.uint8)
        pos = np.zeros((B, n, 2), dtype=np.int32)
        hp = np.zeros((B, n), dtype=np.int32)
        flag = np.zeros((B, n), dtype=np.uint8)
        inv = np.zeros((B, n), dtype=np.int32)
        sc = np.zeros((B, 5), dtype=np.int32)
        stats = np.zeros((B, N_METRICS, n), dtype=np.uint32)
        visits = np.zeros((B, n, g, g), dtype=np.uint8)
        self.L.ctf_oracle_batch_get_state(self._h, _p(grid), _p(pos), _p(hp), _p(flag), _p(inv), _p(sc), _p(stats), _p(visits))
        hp_f = np.zeros((B, n), dtype=np.float64)
        self.L.ctf_oracle_batch_get_hp_f(self._h, _p(hp_f))
        return {
            "grid": grid,
            "pos": pos.astype(np.uint8),
            "hp_q": hp,
            "hp": hp_f,
            "has_flag": flag,
            "inventory": inv,
            "step": sc[:, 0].astype(np.int64),
            "episode": sc[:, 1].astype(np.int64),
            "captures": sc[:, 2:4].astype(np.int64),
            "done": sc[:, 4].astype(bool),
            "stats": stats.astype(np.int64),
            "visits": visits,
        }

    def set_state(self, grid, pos, hp_q, has_flag, inventory, step, episode, captures, hp=None):
        B = self.B
        if hp is not None:   # float HP (cfg.hp_float)
            self.L.ctf_oracle_batch_set_hp_f(self._h, _p(np.ascontiguousarray(np.asarray(hp, dtype=np.float64).reshape(B, self.N))))
        grid = np.ascontiguousarray(np.asarray(grid, dtype=np.uint8))
        pos = np.ascontiguousarray(np.asarray(pos, dtype=np.int32))
        hp = np.ascontiguousarray(np.asarray(hp_q, dtype=np.int32))
        flag = np.ascontiguousarray(np.asarray(has_flag, dtype=np.uint8))
        inv = np.ascontiguousarray(np.asarray(inventory, dtype=np.int32))
        sc = np.zeros((B, 5), dtype=np.int32)
        sc[:, 0] = np.asarray(step).reshape(B)
        sc[:, 1] = np.asarray(episode).reshape(B)
        sc[:, 2:4] = np.asarray(captures).reshape(B, 2)
        sc[:, 4] = sc[:, 0] >= self.ce.GAME_STEPS
        self.L.ctf_oracle_batch_set_state(self._h, _p(grid), _p(pos), _p(hp), _p(flag), _p(inv), _p(sc))


def run_baseline(compiled: CompiledEnv, n_envs: int, steps: int, seed: int, with_obs: bool, n_threads: int) -> float:
    """Multi-threaded CPU run of the oracle (bench.py cpu_baseline / --impl reference). Returns a checksum."""
    return float(lib().ctf_oracle_run_baseline(C.byref(compiled.cfg), n_envs, steps, seed, int(with_obs), n_threads))


def f64_to_f16_to_f32(x: float) -> float:
    return float(lib().ctf_oracle_f64_to_f16_to_f32(float(x)))
