"""The C-ABI library loads and exports every symbol include/ctf_b200.h declares (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

from helpers import compiled
from marl_ctf_development_b200 import _native
from marl_ctf_development_b200.build import build_native
from marl_ctf_development_b200.config import CtfConfig

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    build_native()  # no-op when the in-tree .so is current
    return _native.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "ctf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctf_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree_on_the_symbol_list():
    assert _declared_functions() == sorted(_native.EXPORTS)


def test_every_declared_symbol_is_exported(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), name


def test_config_layout_matches(lib):
    assert lib.ctf_config_size() == C.sizeof(CtfConfig)
    assert lib.ctf_abi_version() == _native.ABI_VERSION


def test_create_fails_loudly_without_a_device_or_with_bad_arguments(lib):
    import torch

    ce = compiled("8_arena")
    h = C.c_void_p()
    if not torch.cuda.is_available():
        rc = lib.ctf_create(C.byref(ce.cfg), 4, 0, 0, 0, 0, 0, C.byref(h))
        assert rc == -3 and b"no CPU fallback" in lib.ctf_last_error()
    rc = lib.ctf_create(C.byref(ce.cfg), 0, 0, 0, 0, 0, 0, C.byref(h))
    assert rc == -1 and b"num_envs" in lib.ctf_last_error()
    rc = lib.ctf_create(C.byref(ce.cfg), 4, 0, 0, 0, 7, 0, C.byref(h))
    assert rc == -1


def test_python_env_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from marl_ctf_development_b200 import GridworldCtfGPU, experiment_env_config

    with pytest.raises(_native.NativeError):
        GridworldCtfGPU(**experiment_env_config("8_arena"), num_envs=2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "marl_ctf_development_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "libctf_oracle" not in text and "ctf_oracle_" not in text, f"{f} links the oracle"
