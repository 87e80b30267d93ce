"""ctypes binding of libctf_b200.so (include/ctf_b200.h).  Fails loudly: no library, no env."""
from __future__ import annotations

import ctypes as C
import os

from .config import CtfConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
# CTF_B200_LIB selects another build of the same extension (kernel A/B experiments); default is the in-tree one
LIB_PATH = os.environ.get("CTF_B200_LIB") or os.path.join(_HERE, "libctf_b200.so")

# every symbol include/ctf_b200.h declares
EXPORTS = (
    "ctf_last_error",
    "ctf_abi_version",
    "ctf_config_size",
    "ctf_create",
    "ctf_destroy",
    "ctf_get_sizes",
    "ctf_reset",
    "ctf_step",
    "ctf_observe",
    "ctf_unpack_obs",
    "ctf_stats_sum",
    "ctf_take_faults",
    "ctf_step_host",
    "ctf_get_kernel_info",
)
ABI_VERSION = 4
FAULT_BAD_ACTION, FAULT_RESPAWN_BLOCKED = 1, 2


class CtfState(C.Structure):
    _fields_ = [
        ("grid", C.c_void_p),
        ("agents", C.c_void_p),
        ("envs", C.c_void_p),
        ("stats", C.c_void_p),
        ("visits", C.c_void_p),
        ("hp", C.c_void_p),
    ]


class CtfOutputs(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("obs_bits", C.c_void_p), ("meta", C.c_void_p), ("rewards", C.c_void_p), ("dones", C.c_void_p)]


class CtfSizes(C.Structure):
    _fields_ = [
        (name, C.c_size_t)
        for name in (
            "grid_stride",
            "grid_bytes",
            "agents_bytes",
            "envs_bytes",
            "stats_bytes",
            "visits_bytes",
            "obs_bytes",
            "meta_bytes",
            "rewards_bytes",
            "dones_bytes",
            "obs_elems_per_env",
            "meta_elems_per_env",
            "obs_bits_bytes",
            "bits_words_per_agent",
            "hp_bytes",
        )
    ]


class CtfKernelInfo(C.Structure):
    _fields_ = [("persistent", C.c_int), ("logic_warps", C.c_int), ("stream_warps", C.c_int), ("ctas", C.c_int),
                ("min_envs_for_persistent", C.c_int64), ("warp_per_env_ctas_per_sm", C.c_int), ("reserved", C.c_int)]


class NativeError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the extension; raises if it is missing (build with ``python -m marl_ctf_development_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("CTF_B200_LIB"):
        # a fresh checkout has no .so (built artefacts are git-ignored): compile it in-tree with nvcc
        from .build import build_native

        try:
            build_native()
        except Exception as exc:
            if not os.path.exists(LIB_PATH):
                raise NativeError(
                    f"{LIB_PATH} is missing and could not be built ({exc}). "
                    "The step path is CUDA-only: there is no CPU fallback."
                ) from exc
    if not os.path.exists(LIB_PATH):
        raise NativeError(f"{LIB_PATH} not found. The step path is CUDA-only: there is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for sym in EXPORTS:
        if not hasattr(L, sym):
            raise NativeError(f"{LIB_PATH} does not export {sym}")
    L.ctf_last_error.restype = C.c_char_p
    L.ctf_abi_version.restype = C.c_int
    L.ctf_config_size.restype = C.c_size_t
    L.ctf_create.argtypes = [C.POINTER(CtfConfig), C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.ctf_destroy.argtypes = [C.c_void_p]
    L.ctf_get_sizes.argtypes = [C.c_void_p, C.POINTER(CtfSizes)]
    L.ctf_reset.argtypes = [C.c_void_p, CtfState, CtfOutputs, C.c_int, C.c_void_p]
    L.ctf_step.argtypes = [C.c_void_p, CtfState, C.c_void_p, CtfOutputs, C.c_void_p]
    L.ctf_observe.argtypes = [C.c_void_p, CtfState, C.c_void_p, CtfOutputs, C.c_void_p]
    L.ctf_unpack_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
    L.ctf_stats_sum.argtypes = [C.c_void_p, CtfState, C.c_void_p, C.c_void_p]
    L.ctf_take_faults.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    L.ctf_get_kernel_info.argtypes = [C.c_void_p, C.POINTER(CtfKernelInfo)]
    L.ctf_step_host.argtypes = [C.c_void_p, CtfState, C.c_void_p, CtfOutputs, C.c_void_p, C.c_void_p, C.c_void_p]
    if L.ctf_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI version mismatch: library {L.ctf_abi_version()}, binding {ABI_VERSION}")
    if L.ctf_config_size() != C.sizeof(CtfConfig):
        raise NativeError(f"ctf_config_t layout mismatch: library {L.ctf_config_size()} B, binding {C.sizeof(CtfConfig)} B")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise NativeError(f"ctf_b200 error {rc}: {load().ctf_last_error().decode()}")
