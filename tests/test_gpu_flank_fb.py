"""The packed forward / backward flank kernels (k_flank_fwd + k_flank_bwd → dp_flank_fwd / dp_flank_bwd / fb_finish) against the oracle's populate with a flank state:
the path hmm::evaluate takes for every candidate near a haplotype flank (pair_hmm.hpp:743-764: traceback DP + calculate_flank_score).
Shapes chosen so that the kernel's own corners are hit: both flanks inside one window, boundaries inside the first / last 2B columns,
co-optimal paths that cross a boundary at different cells (repeat-rich haplotypes: the kernel hands those to the labelled DP),
reads shorter than 2*band (routed to the labelled DP by the classify pass), haplotypes of different lengths (a lane's two windows
with different boundary columns)."""
import numpy as np
import pytest

from helpers import ACGT

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def _compare(engine, coracle, band_req, haps, reads, flanks, mapit=False, dp_only=True):
    from octopus_b200 import HaplotypeLikelihoodModel
    band = HaplotypeLikelihoodModel(HaplotypeLikelihoodModel.Config(max_indel_error=band_req)).pad_requirement()
    cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band_req, use_mapping_quality=False, disable_naive_shortcut=dp_only, map_positions=mapit)
    rc, want, wst = coracle.populate(band, haps, reads, None, flanks, use_mapping_quality=False, dp_only=dp_only, map_positions=mapit)
    got, st = engine.populate(cfg, haps, reads, None, flanks, want_status=True)
    ok = wst == 0
    assert np.array_equal(st[~ok], wst[~ok])
    bad = np.argwhere(got[ok] != want[ok])
    assert np.array_equal(got[ok], want[ok]), (band_req, flanks, len(bad), got[ok][got[ok] != want[ok]][:4], want[ok][got[ok] != want[ok]][:4])   # -ln10/10 * integer: exact


def _repeat_region(rng, hap_len, n_haps, n_reads, read_lens, band):
    """Low-complexity haplotypes (homopolymers and short tandem repeats everywhere, constant gap penalties inside them) and reads
    with indels inside the repeats: many candidates have co-optimal alignments that place the indel on either side of a flank boundary."""
    from octopus_b200.batch import pack_haplotypes, pack_reads
    units = [b"A", b"T", b"AC", b"AG", b"CAG", b"TTA", b"G", b"GT"]
    parts = []
    while sum(len(p) for p in parts) < hap_len + 12:
        u = units[int(rng.integers(0, len(units)))]
        parts.append(u * int(rng.integers(2, 9)))
        if rng.random() < 0.3:
            parts.append(bytes(ACGT[rng.integers(0, 4, int(rng.integers(1, 5)))]))
    base = np.frombuffer(b"".join(parts), dtype=np.uint8)[:hap_len + 12].copy()
    seqs, mf, pf, mr, pr, go, ge = [], [], [], [], [], [], []
    for h in range(n_haps):
        s = base.copy()
        for _ in range(int(rng.integers(0, 3))):
            p = int(rng.integers(2, len(s) - 6)); k = int(rng.integers(1, 4))
            s = np.concatenate([s[:p], s[p + k:], ACGT[rng.integers(0, 4, k)]]) if rng.random() < 0.5 else np.concatenate([s[:p], s[p - k:p], s[p:]])[:len(base)]
        s = s[:hap_len - (h % 3)]                        # different haplotype lengths: different right-flank columns per window
        seqs.append(s)
        mf.append(np.roll(s, 1)); mr.append(np.roll(s, -1))
        pf.append(np.full(len(s), 40, np.int8)); pr.append(np.full(len(s), 40, np.int8))
        go.append(np.full(len(s), 20, np.int8)); ge.append(np.full(len(s), 3, np.int8))     # constant penalties: ties by construction
    haps = pack_haplotypes(seqs, mf, pf, mr, pr, go, ge, begin=np.zeros(n_haps, dtype=np.int64))
    bases, quals, begin = [], [], []
    for r in range(n_reads):
        L = int(rng.choice(read_lens))
        src = seqs[int(rng.integers(0, n_haps))]
        p = int(rng.integers(band, max(band + 1, len(src) - L - band)))
        b = src[p:p + L + 4].copy()
        u = rng.random()
        if u < 0.4 and L > 8:
            i = int(rng.integers(2, L - 2)); b = np.concatenate([b[:i], b[i + 1:]])       # deletion in the read
        elif u < 0.8 and L > 8:
            i = int(rng.integers(2, L - 2)); b = np.concatenate([b[:i], b[i:i + 1], b[i:]])   # duplicated base: insertion inside a run
        b = b[:L]
        if len(b) < L:
            b = np.concatenate([b, ACGT[rng.integers(0, 4, L - len(b))]])
        bases.append(b)
        quals.append(np.full(L, 30, np.uint8) if r % 2 else rng.integers(2, 42, L).astype(np.uint8))
        begin.append(p)
    reads = pack_reads(bases, quals, mapq=np.full(n_reads, 60, np.uint8), reverse=(rng.random(n_reads) < 0.5).astype(np.uint8),
                       begin=np.asarray(begin, dtype=np.int64))
    return haps, reads


@pytest.mark.parametrize("band_req", [8, 16, 32])
def test_flank_state_on_a_benchmark_shaped_region(engine, coracle, band_req):
    """The bench's `--flank` shape: 150-base reads over 300-base haplotypes; every candidate window holds one or both flank boundaries."""
    from octopus_b200 import synth
    rng = np.random.default_rng(900 + band_req)
    haps = synth.make_haplotypes(rng, 40, 300)
    reads = synth.make_reads(rng, haps, 700, (150,), band_req, "q30")
    for flanks in ((60, 60), (100, 100), (0, 90), (130, 0), (20, 10)):
        _compare(engine, coracle, band_req, haps, reads, flanks)


@pytest.mark.parametrize("band_req", [8, 16, 32])
def test_flank_state_with_co_optimal_paths_across_the_boundary(engine, coracle, band_req):
    rng = np.random.default_rng(950 + band_req)
    for trial in range(3):
        haps, reads = _repeat_region(rng, hap_len=240, n_haps=24, n_reads=300, read_lens=[2 * band_req - 3, 2 * band_req, 76, 100, 150][(1 if band_req == 32 else 0):], band=band_req)
        flanks = (int(rng.integers(20, 110)), int(rng.integers(20, 110)))
        _compare(engine, coracle, band_req, haps, reads, flanks)


def test_flank_state_with_mapped_candidate_positions(engine, coracle):
    """Production semantics: naive shortcut on, device k-mer mapper on — several candidate windows per pair, each with its own boundaries."""
    rng = np.random.default_rng(977)
    haps, reads = _repeat_region(rng, hap_len=300, n_haps=16, n_reads=200, read_lens=[60, 100, 150], band=16)
    for flanks in ((70, 80), (150, 40)):
        _compare(engine, coracle, 16, haps, reads, flanks, mapit=True, dp_only=False)


@pytest.mark.parametrize("band_req", [8, 16])
def test_flank_state_over_many_regions_with_few_haplotypes(engine, coracle, band_req):
    """The call shape Octopus really has (one small region after the other, a handful of haplotypes each): phmm_populate_regions cuts a
    warp of the flank kernels into lane groups, each with its own read (different lengths, regions and flank states in one warp)."""
    from octopus_b200 import HaplotypeLikelihoodModel
    from octopus_b200.batch import concat_blocks
    from helpers import random_region
    rng = np.random.default_rng(990 + band_req)
    hap_blocks, read_blocks, flanks = [], [], []
    for g in range(14):
        h, r = random_region(rng, band_req, n_haps=int(rng.choice([1, 2, 3, 5, 8])), n_reads=int(rng.integers(3, 60)),
                             hap_len=2 * band_req + int(rng.choice([180, 240, 300])), read_len_choices=[2 * band_req, 50, 76, 101, 150],
                             read_n_rate=0.03, edge_reads=(g % 2 == 0))
        hap_blocks.append(h); read_blocks.append(r)
        flanks.append((int(rng.integers(0, 100)), int(rng.integers(0, 100))) if g % 4 else None)
    haps, reads, hf, rf = concat_blocks(hap_blocks, read_blocks)
    for mapit, dp_only in ((False, True), (True, False)):
        cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band_req, use_mapping_quality=False, disable_naive_shortcut=dp_only, map_positions=mapit)
        flat, off, st = engine.populate_regions(cfg, haps, reads, hf, rf, flank_states=flanks, want_status=True)
        mats = engine.split_regions(flat, off, hf, rf)
        sts = engine.split_regions(st, off, hf, rf)
        for g, (h, r) in enumerate(zip(hap_blocks, read_blocks)):
            rc, want, wst = coracle.populate(band_req, h, r, None, flanks[g], use_mapping_quality=False, dp_only=dp_only, map_positions=mapit)
            ok = wst == 0
            assert np.array_equal(sts[g][~ok], wst[~ok]), (band_req, g)
            assert np.array_equal(mats[g][ok], want[ok]), (band_req, mapit, g)
