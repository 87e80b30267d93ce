"""Two ranks on two GPUs (skipped on a single-GPU box): sharded envs + NCCL all-reduce == one oracle batch."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import compiled
from oracle.ctf_oracle import OracleBatch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_shards_and_nccl_statistics(tmp_path):
    B, steps, seed, world = 96, 150, 12, 2
    port = 30000 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py"), str(tmp_path), str(B), str(steps), str(seed)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-3000:]
    ce = compiled("8_arena")
    whole = OracleBatch(ce, world * B, seed=seed, env_id_base=0)
    acts = np.random.default_rng(3).integers(0, 9, (steps, world * B, ce.N_AGENTS)).astype(np.uint8)
    for t in range(steps):
        whole.step(acts[t])
    so = whole.state()
    want = so["stats"].sum(0)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"total_{r}.npy"), want)              # every rank holds the global sum
        assert np.array_equal(np.load(tmp_path / f"grid_{r}.npy"), so["grid"][r * B : (r + 1) * B])
        caps = json.load(open(tmp_path / f"caps_{r}.json"))
        assert caps["team_tag_count"] == {"0": int(want[0, 0::2].sum()), "1": int(want[0, 1::2].sum())}
