"""Strided / peer output: phmm_populate_ld and the CUDA IPC mapping (octopus_b200/peer.py) — every rank's epilogue kernel writes its
columns of the [H, R_total] matrix into the owner's memory. Two PROCESSES on one GPU here (CUDA IPC works within a device); the
2- and 8-GPU runs of bench.py check the same through their ``gather_check`` digest."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import random_region

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_populate_into_a_column_window_equals_the_dense_call(engine):
    import torch
    from octopus_b200 import HaplotypeLikelihoodModel
    rng = np.random.default_rng(5)
    for band, flanks, mapit in ((16, None, False), (16, (20, 30), True), (64, None, False)):
        haps, reads = random_region(rng, band, n_haps=23, n_reads=511, hap_len=2 * band + 280, read_len_choices=[76, 100, 150], read_n_rate=0.03)
        cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band, disable_naive_shortcut=not mapit, map_positions=mapit)
        dh, dr = haps.to_device("cuda:0"), reads.to_device("cuda:0")
        dense, st = engine.populate(cfg, dh, dr, None, flanks, want_status=True)
        big = torch.full((haps.n, reads.n + 700), 7.0, dtype=torch.float64, device="cuda:0")
        got, st2 = engine.populate(cfg, dh, dr, None, flanks, out=big[:, 300:300 + reads.n], want_status=True)
        assert torch.equal(big[:, 300:300 + reads.n], dense) and torch.equal(st, st2)
        assert bool((big[:, :300] == 7.0).all()) and bool((big[:, 300 + reads.n:] == 7.0).all())       # nothing outside the window is touched
    with pytest.raises(ValueError):
        engine.populate(cfg, dh, dr, out=torch.empty((reads.n, haps.n), dtype=torch.float64, device="cuda:0").t())


_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import random_region
from octopus_b200 import HaplotypeLikelihoodModel, PairHMMEngine, shard
from octopus_b200.peer import PeerBuffer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
rng = np.random.default_rng(11)                       # the same batch in every process
band = 16
haps, reads = random_region(rng, band, n_haps=37, n_reads=1001, hap_len=330, read_len_choices=[76, 100, 150], read_n_rate=0.02)
cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band)
eng = PairHMMEngine(0)
H, R = haps.n, reads.n
buf = PeerBuffer(2 * H * R * 8, 0, rank, world)       # slot 0: the read-sharded matrix; slot 1: a stack of per-rank matrices
mine, (lo, hi) = shard.shard_reads(reads, world, rank)
dh, dr = haps.to_device("cuda:0"), mine.to_device("cuda:0")
eng.populate(cfg, dh, dr, flank_state=(25, 40), out=buf.view(lo * 8, H, hi - lo, ld=R))
eng.populate(cfg, dh, dr, out=buf.view(H * R * 8 + rank * H * (R // world + 1) * 8, H, hi - lo))
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    whole = eng.populate(cfg, haps.to_device("cuda:0"), reads.to_device("cuda:0"), flank_state=(25, 40))
    assert torch.equal(buf.owner_tensor(0, (H, R)), whole), "peer-written matrix differs from the one-process result"
    plain = eng.populate(cfg, haps.to_device("cuda:0"), reads.to_device("cuda:0"))
    for k in range(world):
        a, b = shard.split_range(R, world, k)
        slab = buf.owner_tensor(H * R * 8 + k * H * (R // world + 1) * 8, (H, b - a))
        assert torch.equal(slab, plain[:, a:b]), k
    print("PEER_OK")
dist.barrier()
buf.close()
dist.destroy_process_group()
"""


def test_two_processes_store_into_one_owner_matrix(tmp_path):
    script = tmp_path / "peer_worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29647")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "PEER_OK" in outs[0][0]
