"""Deterministic synthetic (reads x haplotypes) workloads of BASELINE.json's shapes.

Haplotypes: a uniform random ACGT "reference" per region with per-haplotype SNVs (~1/100 bp) and short indels
(~1/150 bp), cut to exactly ``hap_len``; 1 % of haplotypes carry one 'N'. Penalties are drawn from the value range of
the reference's error-model tables (core/models/error/error_model_factory.cpp:220-523): gap_open 3..45,
gap_extend 1..10 capped at the position's gap_open (no short-read error model makes an extension dearer than the opening,
tests/test_error_model.py; ``ordered_penalties=False`` drops the cap and the DP kernels then run their general deletion
update, phmm_device.cuh dp_pair OGE), snv_prior 1..125; snv_mask is the neighbouring base
(repeat_based_snv_error_model.cpp:174-178).
Reads: source haplotype uniform, start uniform in [band, hap_len - L - band] (the in-range rule,
haplotype_likelihood_model.cpp:187-207), substitutions at 10^(-q/10), one indel with probability ~2e-3 * L,
strand 50/50, mapq 60. All generation is vectorised numpy; seeds are fixed per config.
"""
import numpy as np

from .batch import HaplotypeBlock, ReadBlock

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

CONFIGS = {
    # name: reads, haplotypes, read lengths, hap_len, band, quality profile   (BASELINE.json "configs")
    "C1": dict(n_reads=1000, n_haps=8, read_lens=(150,), hap_len=300, band=8, quals="q30", seed=0xC0FFEE + 1),
    "C2": dict(n_reads=100_000, n_haps=64, read_lens=(150,), hap_len=300, band=16, quals="q30", seed=0xC0FFEE + 2),
    "C3": dict(n_reads=1_000_000, n_haps=128, read_lens=(150,), hap_len=300, band=16, quals="empirical", seed=0xC0FFEE + 3),
    "C4": dict(n_reads=500_000, n_haps=256, read_lens=(76, 150, 250), hap_len=500, band=32, quals="empirical", seed=0xC0FFEE + 4),
    "C5": dict(n_reads=10_000, n_haps=1024, read_lens=(150,), hap_len=300, band=16, quals="empirical", seed=0xC0FFEE + 5),
}


def make_haplotypes(rng, n_haps, hap_len, ordered_penalties=True):
    slack = 16
    base = _ACGT[rng.integers(0, 4, hap_len + slack)]
    seqs = np.empty((n_haps, hap_len), dtype=np.uint8)
    for h in range(n_haps):
        s = base.copy()
        n_snv = rng.poisson(hap_len / 100.0)
        if n_snv:
            idx = rng.integers(0, len(s), n_snv)
            s[idx] = _ACGT[rng.integers(0, 4, n_snv)]
        n_indel = rng.poisson(hap_len / 150.0)
        for _ in range(n_indel):
            p = int(rng.integers(1, len(s) - 4))
            k = int(rng.integers(1, 4))
            if rng.random() < 0.5 and len(s) - k >= hap_len:
                s = np.concatenate([s[:p], s[p + k:]])
            else:
                s = np.concatenate([s[:p], _ACGT[rng.integers(0, 4, k)], s[p:]])
        seqs[h] = s[:hap_len]
        if rng.random() < 0.01:
            seqs[h, int(rng.integers(0, hap_len))] = ord("N")
    total = n_haps * hap_len
    flat = seqs.reshape(-1)
    mask_f = np.roll(seqs, 1, axis=1).reshape(-1).copy()     # seq[i-1]
    mask_r = np.roll(seqs, -1, axis=1).reshape(-1).copy()    # seq[i+1]
    off = np.arange(n_haps + 1, dtype=np.int64) * hap_len
    prior_f, prior_r = rng.integers(1, 126, total).astype(np.int8), rng.integers(1, 126, total).astype(np.int8)
    gap_open, gap_extend = rng.integers(3, 46, total).astype(np.int8), rng.integers(1, 11, total).astype(np.int8)
    if ordered_penalties:
        gap_extend = np.minimum(gap_extend, gap_open)
    return HaplotypeBlock(off, flat.copy(), mask_f, prior_f, mask_r, prior_r, gap_open, gap_extend, np.zeros(n_haps, dtype=np.int64))


def _qualities(rng, n, L, profile):
    if profile == "q30":
        return np.full((n, L), 30, dtype=np.uint8)
    # "empirical": Illumina-binned {2: 2 %, 12: 3 %, 23: 10 %, 37: 85 %}, low bins concentrated in the last 20 % of the read
    q = np.full((n, L), 37, dtype=np.uint8)
    u = rng.random((n, L), dtype=np.float32)
    tail = np.arange(L) >= int(0.8 * L)
    scale = np.where(tail, 3.0, 0.5).astype(np.float32)      # overall mass ≈ 15 % low-quality, mostly in the tail
    q[u < 0.15 * scale] = 23
    q[u < 0.05 * scale] = 12
    q[u < 0.02 * scale] = 2
    return q


def make_reads(rng, haps: HaplotypeBlock, n_reads, read_lens, band, profile):
    hap_len = haps.length(0)
    H = haps.n
    seqs = haps.seq.reshape(H, hap_len)
    lens = np.asarray(read_lens, dtype=np.int64)
    which = rng.integers(0, len(lens), n_reads) if len(lens) > 1 else np.zeros(n_reads, dtype=np.int64)
    L_of = lens[which]
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(L_of, out=off[1:])
    bases = np.empty(int(off[-1]), dtype=np.uint8)
    quals = np.empty(int(off[-1]), dtype=np.uint8)
    begin = np.empty(n_reads, dtype=np.int64)
    p_err = np.power(10.0, -np.arange(256, dtype=np.float64) / 10.0).astype(np.float32)   # substitution probability by quality
    chunk = 100_000                                                                       # keeps the temporaries cache-sized
    work = []
    for li, L in enumerate(lens):
        idx_all = np.nonzero(which == li)[0]
        for c0 in range(0, len(idx_all), chunk):
            work.append((int(L), idx_all[c0:c0 + chunk]))
    for L, sel in work:
        n = len(sel)
        if n == 0:
            continue
        lo, hi = band, hap_len - L - band
        assert hi >= lo, "haplotype too short for this read length / band"
        src = rng.integers(0, H, n)
        start = rng.integers(lo, hi + 1, n)
        col = np.arange(L)
        idx = start[:, None] + col[None, :]
        # one indel per read with probability 2e-3 * L (deletion: skip a haplotype base; insertion: repeat then overwrite)
        ev = rng.random(n) < 2e-3 * L
        ipos = rng.integers(1, L - 1, n) if L > 2 else np.zeros(n, dtype=np.int64)
        is_del = rng.random(n) < 0.5
        shift = np.where((ev & is_del)[:, None] & (col[None, :] >= ipos[:, None]), 1, 0) \
            - np.where((ev & ~is_del)[:, None] & (col[None, :] > ipos[:, None]), 1, 0)
        idx = np.clip(idx + shift, 0, hap_len - 1)
        rb = seqs[src[:, None], idx]
        ins = ev & ~is_del
        if ins.any():
            rb[np.nonzero(ins)[0], ipos[ins]] = _ACGT[rng.integers(0, 4, int(ins.sum()))]
        q = _qualities(rng, n, L, profile)
        err = rng.random((n, L), dtype=np.float32) < p_err[q]
        rb = np.where(err, _ACGT[rng.integers(0, 4, (n, L), dtype=np.uint8)], rb)
        rb = np.where(rb == ord("N"), ord("A"), rb).astype(np.uint8)
        dest = off[sel][:, None] + col[None, :]
        bases[dest] = rb
        quals[dest] = q
        begin[sel] = start
    reverse = (rng.random(n_reads) < 0.5).astype(np.uint8)
    return ReadBlock(off, bases, quals, np.full(n_reads, 60, dtype=np.uint8), reverse, begin)


def make_batch(config="C1", n_reads=None, n_haps=None, seed=None, band=None, hap_len=None, read_lens=None, ordered_penalties=True):
    """(haplotypes, reads, band) for a named BASELINE config, optionally down-sized (same shapes, fewer reads / haplotypes)
    or re-shaped (band / haplotype length / read lengths overridden: the wide-band and long-read diagnostics)."""
    c = dict(CONFIGS[config])
    if band is not None:
        c["band"] = int(band)
    if hap_len is not None:
        c["hap_len"] = int(hap_len)
    if read_lens is not None:
        c["read_lens"] = tuple(int(x) for x in read_lens)
    if n_reads is not None:
        c["n_reads"] = int(n_reads)
    if n_haps is not None:
        c["n_haps"] = int(n_haps)
    rng = np.random.default_rng(c["seed"] if seed is None else seed)
    haps = make_haplotypes(rng, c["n_haps"], c["hap_len"], ordered_penalties)
    reads = make_reads(rng, haps, c["n_reads"], c["read_lens"], c["band"], c["quals"])
    return haps, reads, c["band"]


def cells_per_alignment(L, band):
    """Banded DP cells the reference loops over per alignment: 2 * (L + band) * band (simd_pair_hmm.hpp:271)."""
    return 2 * (L + band) * band


def total_cells(haps: HaplotypeBlock, reads: ReadBlock, band):
    lens = np.diff(np.asarray(reads.off if not hasattr(reads.off, "cpu") else reads.off.cpu().numpy()))
    return int((2 * (lens + band) * band).sum()) * haps.n
