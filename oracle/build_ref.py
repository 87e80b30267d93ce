"""TEST INFRASTRUCTURE — byte-compiles the reference's own implementation of the path into oracle/_ref/.

    python oracle/build_ref.py          (run where /root/reference exists; __graft_entry__.build() calls it)

The reference is pure Python: "compiling it from the sources where they lie" means ``py_compile`` of the files on
the step path and of its callers, straight from /root/reference into oracle/_ref/<name>.refc (the .pyc byte stream
under another suffix: the box snapshot drops *.pyc as cache files).  Only build outputs land there — no source text,
no assets; oracle/_ref/ is git-ignored (never in history) but travels to the GPU box like the built .so files.  oracle/ref_shim.py imports the unmodified reference from there when /root/reference is absent, so

* bench.py's ``cpu_baseline`` / ``--impl reference`` can time the REAL reference on the GPU box's host cores
  (kind "reference"), and
* ``-m gpu`` tests can run the reference's own callers (utils.duel, PPOTrainer.get_single_rollout,
  MetricsLogger.harvest_metrics) against the CUDA backend.

Byte code is tied to the interpreter version; the GPU box runs this same image.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("CTF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SUFFIX = ".refc"

MODULES = ("gridworld_ctf", "scenarios", "utils", "ppo", "agent_network", "metrics_logger", "league_training")
EXPERIMENTS = ("0_the_split", "1_fence", "2_jailbreak", "3_one_way_out", "4_keyhole", "5_skittles", "6_the_wall",
               "7_gridlocked", "8_arena")


def build(verbose: bool = True) -> str | None:
    if not os.path.isfile(os.path.join(SRC_ROOT, "gridworld_ctf.py")):
        if verbose:
            print(f"oracle/_ref: {SRC_ROOT} not present, nothing to compile (prebuilt files, if any, are kept)")
        return None
    os.makedirs(OUT, exist_ok=True)
    for stale in os.listdir(OUT):
        if stale.endswith((".pyc", SUFFIX)):
            os.unlink(os.path.join(OUT, stale))
    for name in MODULES + EXPERIMENTS:
        py_compile.compile(os.path.join(SRC_ROOT, name + ".py"), cfile=os.path.join(OUT, name + SUFFIX), doraise=True,
                           optimize=0)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write(f"py_compile of {len(MODULES) + len(EXPERIMENTS)} files from {SRC_ROOT} with Python {sys.version.split()[0]}\n")
    if verbose:
        print(f"built {OUT} ({len(MODULES) + len(EXPERIMENTS)} byte-compiled files)")
    return OUT


if __name__ == "__main__":
    build()
