import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_batches_are_deterministic_and_in_range():
    from octopus_b200 import synth
    h1, r1, band = synth.make_batch("C1")
    h2, r2, _ = synth.make_batch("C1")
    assert band == 8 and h1.n == 8 and r1.n == 1000
    for f in h1._fields:
        assert np.array_equal(getattr(h1, f), getattr(h2, f))
    for f in r1._fields:
        assert np.array_equal(getattr(r1, f), getattr(r2, f))
    L = np.diff(r1.off)
    assert (L == 150).all()
    # the in-range rule of haplotype_likelihood_model.cpp:187-207 holds for every read against every haplotype
    assert (r1.begin >= band).all() and (r1.begin + L + band <= 300).all()
    assert set(np.unique(r1.bases)) <= set(b"ACGT")
    h4, r4, b4 = synth.make_batch("C4", n_reads=3000, n_haps=4)
    assert b4 == 32 and set(np.unique(np.diff(r4.off))) == {76, 150, 250}
    assert synth.total_cells(h1, r1, band) == 8 * 1000 * 2 * (150 + 8) * 8


def test_split_and_position_sharding():
    from helpers import random_positions, random_region
    from octopus_b200 import shard
    for n, w in [(10, 3), (7, 8), (1000, 8), (0, 2)]:
        parts = [shard.split_range(n, w, k) for k in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    rng = np.random.default_rng(1)
    haps, reads = random_region(rng, 8, 3, 11, 120, [30, 40])
    off, pos = random_positions(rng, haps, reads)
    for world in (2, 3):
        for rank in range(world):
            sub, (lo, hi) = shard.shard_reads(reads, world, rank)
            assert sub.n == hi - lo
            for j in range(sub.n):
                assert np.array_equal(sub.read(j)[0], reads.read(lo + j)[0])
            so, sp = shard.shard_positions((off, pos), haps.n, reads.n, lo, hi)
            for h in range(haps.n):
                for j in range(sub.n):
                    a = pos[off[h * reads.n + lo + j]:off[h * reads.n + lo + j + 1]]
                    b = sp[so[h * sub.n + j]:so[h * sub.n + j + 1]]
                    assert np.array_equal(a, b)


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from octopus_b200 import shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
H, R = 5, 23
full = torch.arange(H * R, dtype=torch.float64).reshape(H, R)
lo, hi = shard.split_range(R, world, rank)
got = shard.gather_likelihoods(full[:, lo:hi].clone(), R, world, rank)
slabs = shard.gather_slabs(full * (rank + 1), world, rank)                    # weak-scaling exchange: one matrix per rank
if rank == 0:
    assert torch.equal(got, full)
    assert len(slabs) == world and all(torch.equal(slabs[k], full * (k + 1)) for k in range(world))
    print("GATHER_OK")
else:
    assert got is None and slabs is None
dist.destroy_process_group()
"""


def test_gather_to_rank0_gloo_world2(tmp_path):
    """The N>1 path (static split + gather to rank 0) on CPU with the gloo backend, world_size 2."""
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0][0]
