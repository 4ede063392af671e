"""Full-size runs (BASELINE.json's C2 shape on one GPU) checked through size-independent properties plus a random
sample of pairs against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_full_size_properties(engine, coracle):
    from octopus_b200 import HaplotypeLikelihoodModel, synth
    from octopus_b200.batch import ReadBlock
    haps, reads, band = synth.make_batch("C2")           # 100k reads x 64 haplotypes, band 16: 6.4e6 alignments
    cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band, disable_naive_shortcut=True, use_mapping_quality=False, map_positions=False)
    m = engine.populate(cfg, haps, reads)
    assert m.shape == (64, 100_000) and np.isfinite(m).all() and (m <= 0).all()
    # 1. determinism
    assert np.array_equal(m, engine.populate(cfg, haps, reads))
    # 2. every value is -ln10/10 * (a non-negative integer phred score)
    c = 0.230258509299404568401799145468436420760110148862877297603
    k = np.rint(-m / c)
    assert np.array_equal(-c * k, m) and (k >= 0).all()
    # 3. permuting the reads permutes the columns (scheduling / pairing must not leak between reads)
    rng = np.random.default_rng(0)
    perm = rng.permutation(reads.n)[:20000]
    L = 150
    idx = (reads.off[perm][:, None] + np.arange(L)[None, :]).reshape(-1)
    sub = ReadBlock(np.arange(len(perm) + 1, dtype=np.int64) * L, reads.bases[idx], reads.quals[idx], reads.mapq[perm], reads.reverse[perm], reads.begin[perm])
    assert np.array_equal(engine.populate(cfg, haps, sub), m[:, perm])
    # 4. a random sample of pairs against the oracle (bit-exact: integer score times a constant)
    sel = rng.choice(reads.n, 300, replace=False)
    idx = (reads.off[sel][:, None] + np.arange(L)[None, :]).reshape(-1)
    sample = ReadBlock(np.arange(len(sel) + 1, dtype=np.int64) * L, reads.bases[idx], reads.quals[idx], reads.mapq[sel], reads.reverse[sel], reads.begin[sel])
    rc, want, _ = coracle.populate(band, haps, sample, use_mapping_quality=False, dp_only=True)
    assert rc == 0 and np.array_equal(m[:, sel], want)
    # 5. a read scored against the haplotype it was copied from without errors costs nothing: spot-check exact substrings
    hs = haps.seq.reshape(64, 300)
    for r in sel[:50]:
        b, _ = reads.read(int(r))
        p = int(reads.begin[r])
        for h in range(64):
            if np.array_equal(hs[h, p:p + L], b):
                assert m[h, r] == 0.0


def test_pipelined_host_path_equals_single_pass():
    """Host-resident batches above a size threshold are cut into read chunks and pipelined over two sub-engines (copies of
    one chunk overlap the kernels of the other). Forced here on a small batch through the PHMM_CHUNK_PAIRS test hook, in a
    fresh process (the threshold is read once), and compared with the unchunked device-resident result."""
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from octopus_b200 import HaplotypeLikelihoodModel, PairHMMEngine, synth
haps, reads, band = synth.make_batch("C4", n_reads=9000, n_haps=20)          # ragged lengths, band 32
eng = PairHMMEngine(0)
for kw in (dict(map_positions=False, disable_naive_shortcut=True), dict(), dict(use_mapping_quality=False)):
    cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band, **kw)
    for flanks in (None, (70, 60)):
        host, st = eng.populate(cfg, haps, reads, flank_state=flanks, want_status=True)          # chunked (env hook)
        dev = eng.populate(cfg, haps.to_device("cuda:0"), reads.to_device("cuda:0"), flank_state=flanks).cpu().numpy()
        assert np.array_equal(host, dev), (kw, flanks)
        assert (st == 0).all()
assert eng.launch_count(total=True) > 100
# error paths must come back from both worker threads (no chunk may be left waiting for the other's DP event)
from octopus_b200.api import PhmmError, ShortHaplotypeError
cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band)
bad = synth.make_batch("C4", n_reads=9000, n_haps=20)[0]
bad.gap_open[5] = -3                                                          # penalty outside [0, 127]
try:
    eng.populate(cfg, bad, reads); raise SystemExit("invalid penalty not reported")
except PhmmError as e:
    assert e.code == -1, e
short_haps = synth.make_haplotypes(np.random.default_rng(1), 20, 260)         # shorter than a 250 bp read + its band-32 pad
try:
    eng.populate(cfg, short_haps, reads); raise SystemExit("short haplotype not reported")
except ShortHaplotypeError:
    pass
host = eng.populate(cfg, haps, reads)                                         # and the engine is still usable afterwards
assert np.array_equal(host, eng.populate(cfg, haps.to_device("cuda:0"), reads.to_device("cuda:0")).cpu().numpy())
print("PIPELINE_OK")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    env = dict(os.environ, PHMM_CHUNK_PAIRS="20000")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PIPELINE_OK" in out.stdout, out.stdout + out.stderr
