"""In-tree build of the CUDA extension (libctf_b200.so) for sm_100a with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "ctf_kernels.cu")
HDR = os.path.join(_HERE, "..", "include", "ctf_b200.h")
LIB = os.path.join(_HERE, "libctf_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in (SRC, HDR))


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/ctf_kernels.cu into libctf_b200.so next to this file. Returns the path.

    Safe when several ranks start at once (torchrun on a fresh checkout): one process at a time holds an flock,
    nvcc writes to a temporary file and the finished library is moved into place atomically, so no rank can dlopen
    a half-written file; the ranks that waited find an up-to-date library and do not compile again."""
    if not force and not is_stale():
        return LIB
    import fcntl
    import tempfile

    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():   # another rank built it while this one waited
                return LIB
            fd, tmp = tempfile.mkstemp(prefix=".libctf_b200.", suffix=".so.tmp", dir=_HERE)
            os.close(fd)
            try:
                cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
                proc = subprocess.run(cmd, capture_output=True, text=True)
                if proc.returncode != 0:
                    raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
                os.chmod(tmp, 0o755)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
            if verbose:
                print(proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
