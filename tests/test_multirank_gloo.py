"""N > 1 host logic on CPU with world_size-2 gloo: env-id sharding, statistics all-reduce into the env.metrics
schema, and bench.py's reference arm under a two-rank launch.  The oracle stands in for the device here."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import compiled
from marl_ctf_development_b200.sharding import all_reduce_stats, env_id_base, shard_bounds

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
B_PER_RANK, STEPS, SEED = 24, 120, 5


def _actions(total, n_agents):
    return np.random.default_rng(3).integers(0, 9, (STEPS, total, n_agents)).astype(np.uint8)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.ctf_oracle import OracleBatch

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ce = compiled("8_arena")
    base = env_id_base(B_PER_RANK)  # rank * B, taken from the process group
    assert base == rank * B_PER_RANK
    shard = OracleBatch(ce, B_PER_RANK, seed=SEED, env_id_base=base)
    acts = _actions(world * B_PER_RANK, ce.N_AGENTS)
    for t in range(STEPS):
        shard.step(acts[t, base : base + B_PER_RANK])
    st = shard.state()
    counters = torch.from_numpy(st["stats"].sum(0).astype(np.int64))
    all_reduce_stats(counters)
    np.save(os.path.join(out_dir, f"counters_{rank}.npy"), counters.numpy())
    np.save(os.path.join(out_dir, f"grid_{rank}.npy"), st["grid"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_one_batch(tmp_path):
    from oracle.ctf_oracle import OracleBatch

    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ce = compiled("8_arena")
    whole = OracleBatch(ce, world * B_PER_RANK, seed=SEED, env_id_base=0)
    acts = _actions(world * B_PER_RANK, ce.N_AGENTS)
    for t in range(STEPS):
        whole.step(acts[t])
    st = whole.state()
    want = st["stats"].sum(0)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"counters_{r}.npy"), want)  # both ranks hold the global sum
        assert np.array_equal(np.load(tmp_path / f"grid_{r}.npy"), st["grid"][r * B_PER_RANK : (r + 1) * B_PER_RANK])


def test_metrics_schema_from_reduced_counters():
    """The reduced [13, N] counters rebuild the reference's env.metrics families (gridworld_ctf.py:425-470)."""
    from marl_ctf_development_b200.config import METRIC_NAMES
    from marl_ctf_development_b200.env import metrics_dict

    ce = compiled("8_arena")
    counters = np.arange(13 * 8).reshape(13, 8)
    m = metrics_dict(ce, counters)
    for k, name in enumerate(METRIC_NAMES):
        assert m["team_" + name][0] == counters[k, 0::2].sum() and m["team_" + name][1] == counters[k, 1::2].sum()
        for i in range(8):
            assert m["agent_" + name][i] == counters[k, i]
            assert m["agent_type_" + name][i % 2][ce.AGENT_TYPES[i]] == counters[k, i]  # one agent per (team, type) in 8_arena
    assert set(m) >= {"team_wins", "agent_visitation_maps", "team_flag_captures", "agent_type_tag_count"}


def test_shard_bounds_cover_everything():
    for total in (7, 64, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_reference_arm_under_two_ranks_prints_one_line():
    """bench.py --impl reference launched like the driver does for N = 2: rank 0 prints the line, rank 1 exits 0."""
    port = 31000 + os.getpid() % 2000
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
        "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3",
    ]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    from oracle import ref_shim as rs

    if rs.available():   # the unmodified Python reference (source tree here, oracle/_ref on the GPU box) is what gets timed
        assert d["cpu_baseline"]["kind"] == "reference" and "unmodified gridworld_ctf.GridworldCtf" in d["cpu_baseline"]["sample"]
