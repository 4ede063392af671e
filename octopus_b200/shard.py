"""Multi-GPU: one process per GPU, static split of the work items, one gather of the result slabs to rank 0.

The path shards embarrassingly: every (read, haplotype, position) alignment is independent, so the
reads of a region are split contiguously over the ranks, every rank holds all haplotypes (a few hundred KB), and there
is NO collective inside the computation. The only exchange is the final gather of the [H, R_rank] ln-likelihood slabs
to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def split_range(n, world, rank):
    """Contiguous static split of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_reads(reads, world, rank):
    """The ReadBlock holding this rank's contiguous share of the reads (host arrays)."""
    from .batch import ReadBlock
    lo, hi = split_range(reads.n, world, rank)
    a, b = int(reads.off[lo]), int(reads.off[hi])
    return ReadBlock(reads.off[lo:hi + 1] - reads.off[lo], reads.bases[a:b], reads.quals[a:b], reads.mapq[lo:hi],
                     reads.reverse[lo:hi], reads.begin[lo:hi]), (lo, hi)


def shard_positions(positions, H, R, lo, hi):
    """Column slice [lo, hi) of a [H][R] CSR of candidate positions."""
    if positions is None:
        return None
    off, pos = positions
    n = hi - lo
    new_off = np.zeros(H * n + 1, dtype=np.int64)
    parts = []
    for h in range(H):
        a, b = int(off[h * R + lo]), int(off[h * R + hi])
        parts.append(pos[a:b])
        new_off[h * n + 1:(h + 1) * n + 1] = new_off[h * n] + (off[h * R + lo + 1:h * R + hi + 1] - off[h * R + lo])
    flat = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int32)
    return new_off, (flat if len(flat) else np.zeros(1, dtype=np.int32))


def gather_likelihoods(local, R_total, world, rank, group=None, buffers=None):
    """Gather the per-rank [H, R_rank] slabs into the [H, R_total] matrix on rank 0 (returns None elsewhere).

    ``local`` is a torch tensor (CUDA with the nccl backend, CPU with gloo); the column ranges are those of ``split_range``.
    Equal shares are gathered straight from ``local``; unequal ones are padded to the largest. ``buffers``: an optional dict the
    caller keeps across calls (receive slabs, output matrix, padding slab) so that a step allocates nothing."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    H = local.shape[0]
    width = (R_total + world - 1) // world
    buffers = buffers if buffers is not None else {}
    equal = R_total % world == 0
    if equal:
        send = local.contiguous()
    else:
        send = buffers.get("send")
        if send is None or send.shape != (H, width) or send.device != local.device:
            send = buffers["send"] = torch.zeros((H, width), dtype=local.dtype, device=local.device)
        send[:, :local.shape[1]] = local
    recv = None
    if rank == 0:
        recv = buffers.get("recv")
        if recv is None or len(recv) != world or recv[0].shape != (H, width) or recv[0].device != local.device:
            recv = buffers["recv"] = [torch.empty((H, width), dtype=local.dtype, device=local.device) for _ in range(world)]
    dist.gather(send, recv, dst=0, group=group)
    if rank != 0:
        return None
    out = buffers.get("out")
    if out is None or out.shape != (H, R_total) or out.device != local.device:
        out = buffers["out"] = torch.empty((H, R_total), dtype=local.dtype, device=local.device)
    for k in range(world):
        lo, hi = split_range(R_total, world, k)
        out[:, lo:hi] = recv[k][:, :hi - lo]
    return out


def gather_slabs(local, world, rank, recv=None, group=None):
    """Gather equally-shaped per-rank result matrices to rank 0 as a list (rank order); returns None on the other ranks.

    This is the exchange of the weak-scaling deployment (every rank calls its own regions: the per-rank matrices belong to
    different haplotype sets and are not columns of one matrix), so nothing is padded or re-assembled: one NCCL gather
    straight into the receive slabs. ``recv`` lets the caller reuse the world x local-shaped receive buffers."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local]
    if rank == 0 and recv is None:
        recv = [torch.empty_like(local) for _ in range(world)]
    dist.gather(local.contiguous(), recv if rank == 0 else None, dst=0, group=group)
    return recv if rank == 0 else None
