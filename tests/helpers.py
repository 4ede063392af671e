"""Seeded random inputs shared by the CPU and GPU parity tests."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_alignment_case(rng, band, L, n_rate=0.2, qmax=41, read_n=False, open_ge_extend=False):
    """One raw-kernel input: truth window (W = L + 2*band - 1) and a read derived from it with subs / indels."""
    W = L + 2 * band - 1
    truth = ACGT[rng.integers(0, 4, W)].copy()
    if rng.random() < n_rate:
        truth[rng.integers(0, W)] = ord("N")
    off = int(rng.integers(0, 2 * band))
    src = truth[off:]
    read, i = [], 0
    while len(read) < L:
        u = rng.random()
        if i >= len(src):
            read.append(int(ACGT[rng.integers(0, 4)]))
        elif u < 0.03:
            read.append(int(ACGT[rng.integers(0, 4)])); i += 1
        elif u < 0.05:
            read.append(int(ACGT[rng.integers(0, 4)]))
        elif u < 0.07:
            i += 1
        else:
            read.append(int(src[i]) if src[i] != ord("N") else ord("A")); i += 1
    read = np.array(read[:L], dtype=np.uint8)
    if read_n and L > 2:
        read[rng.integers(0, L)] = ord("N")
    gap_open = rng.integers(3, 46, W).astype(np.int8)
    gap_extend = rng.integers(1, 11, W).astype(np.int8)
    if open_ge_extend:          # what every built-in error model produces: an extension never costs more than the opening
        gap_extend = np.minimum(gap_extend, gap_open)
    return dict(
        truth=truth, read=read,
        quals=rng.integers(2, qmax + 1, L).astype(np.uint8),
        gap_open=gap_open,
        gap_extend=gap_extend,
        snv_mask=np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, W)].copy(),
        snv_prior=rng.integers(1, 126, W).astype(np.int8),
    )


def random_region(rng, band, n_haps, n_reads, hap_len, read_len_choices, mutate=True, read_n_rate=0.0, edge_reads=True):
    """A small (haplotypes x reads) region as HaplotypeBlock / ReadBlock, with reads that overhang, sit at edges, etc."""
    from octopus_b200.batch import pack_haplotypes, pack_reads
    base = ACGT[rng.integers(0, 4, hap_len + 8)]
    seqs, mf, pf, mr, pr, go, ge = [], [], [], [], [], [], []
    for h in range(n_haps):
        s = base.copy()
        if mutate:
            for _ in range(int(rng.integers(0, 4))):
                s[rng.integers(0, len(s))] = ACGT[rng.integers(0, 4)]
            if rng.random() < 0.5:
                p = int(rng.integers(2, len(s) - 4)); s = np.concatenate([s[:p], s[p + 2:], ACGT[rng.integers(0, 4, 2)]])
            if rng.random() < 0.15:
                s[rng.integers(0, hap_len)] = ord("N")
        s = s[:hap_len]
        seqs.append(s)
        mf.append(np.roll(s, 1)); mr.append(np.roll(s, -1))
        pf.append(rng.integers(1, 126, hap_len).astype(np.int8)); pr.append(rng.integers(1, 126, hap_len).astype(np.int8))
        go.append(rng.integers(3, 46, hap_len).astype(np.int8)); ge.append(rng.integers(1, 11, hap_len).astype(np.int8))
    haps = pack_haplotypes(seqs, mf, pf, mr, pr, go, ge, begin=np.zeros(n_haps, dtype=np.int64))
    bases, quals, begin = [], [], []
    for r in range(n_reads):
        L = int(rng.choice(read_len_choices))
        lo, hi = (0, hap_len - L) if edge_reads else (band, hap_len - L - band)
        p = int(rng.integers(lo, max(lo, hi) + 1))
        src = seqs[int(rng.integers(0, n_haps))]
        b = src[p:p + L].copy()
        if len(b) < L:
            b = np.concatenate([b, ACGT[rng.integers(0, 4, L - len(b))]])
        b[b == ord("N")] = ord("A")
        for _ in range(int(rng.choice([0, 0, 1, 1, 2, 4]))):
            b[rng.integers(0, L)] = ACGT[rng.integers(0, 4)]
        if rng.random() < 0.2 and L > 4:
            i = int(rng.integers(1, L - 2)); b = np.concatenate([b[:i], b[i + 1:], ACGT[rng.integers(0, 4, 1)]])
        if rng.random() < read_n_rate:
            b[rng.integers(0, L)] = ord("N")
        bases.append(b)
        quals.append(rng.integers(2, 42, L).astype(np.uint8))
        begin.append(p)
    reads = pack_reads(bases, quals, mapq=rng.choice([0, 10, 29, 60, 255], n_reads).astype(np.uint8),
                       reverse=(rng.random(n_reads) < 0.5).astype(np.uint8), begin=np.asarray(begin, dtype=np.int64))
    return haps, reads


def random_positions(rng, haps, reads, max_listed=4):
    """Candidate mapping positions per (haplotype, read) pair, near the read's original position, CSR in [H][R] order."""
    from octopus_b200.batch import pack_positions
    lists = []
    for h in range(haps.n):
        row = []
        for r in range(reads.n):
            p0 = int(reads.begin[r])
            n = int(rng.integers(0, max_listed + 1))
            ps = sorted({int(np.clip(p0 + rng.integers(-12, 13), 0, haps.length(h))) for _ in range(n)})
            if ps and rng.random() < 0.5:
                ps[0] = p0
            row.append(sorted(set(ps)))
        lists.append(row)
    return pack_positions(lists, haps.n, reads.n)


def n_rich_flank_region(rng, hap_len=220, n_haps=6, n_reads=60):
    """Haplotypes with several 'N's (flanks included), SNV priors and read qualities that include 0 and 1, reads over the whole
    haplotype, and a flank state: the corner where the reference's flank replay charges a truth-'N' mismatch 2 although its DP
    charged min(q', 2). Returns (haplotypes, reads, (lhs_flank, rhs_flank))."""
    from octopus_b200.batch import pack_haplotypes, pack_reads
    base = ACGT[rng.integers(0, 4, hap_len)]
    seqs, mf, pf, mr, pr, go, ge = [], [], [], [], [], [], []
    for h in range(n_haps):
        s = base.copy()
        for _ in range(int(rng.integers(0, 3))):
            s[rng.integers(0, hap_len)] = ACGT[rng.integers(0, 4)]
        s[rng.integers(0, hap_len, int(rng.integers(3, 12)))] = ord("N")
        seqs.append(s)
        mf.append(ACGT[rng.integers(0, 4, hap_len)].copy()); mr.append(np.roll(s, -1))
        pf.append(rng.choice([0, 1, 2, 30, 125], hap_len).astype(np.int8)); pr.append(rng.choice([0, 1, 2, 30, 125], hap_len).astype(np.int8))
        go.append(rng.integers(3, 46, hap_len).astype(np.int8)); ge.append(rng.integers(1, 11, hap_len).astype(np.int8))
    haps = pack_haplotypes(seqs, mf, pf, mr, pr, go, ge, begin=np.zeros(n_haps, dtype=np.int64))
    bases, quals, begins = [], [], []
    for r in range(n_reads):
        L = int(rng.choice([30, 50, 76]))
        p = int(rng.integers(0, hap_len - L + 1))
        b = base[p:p + L].copy()
        for _ in range(int(rng.integers(0, 3))):
            b[rng.integers(0, L)] = ACGT[rng.integers(0, 4)]
        bases.append(b)
        quals.append((rng.choice([0, 1, 2, 20, 40], L) if r % 2 else rng.integers(2, 42, L)).astype(np.uint8))
        begins.append(p)
    reads = pack_reads(bases, quals, mapq=np.full(n_reads, 60, np.uint8), reverse=(rng.random(n_reads) < 0.5).astype(np.uint8),
                       begin=np.array(begins, dtype=np.int64))
    return haps, reads, (int(rng.integers(40, 90)), int(rng.integers(40, 90)))
