"""HaplotypeLikelihoodArray container semantics (reference: haplotype_likelihood_array.cpp:200-409) — host logic only:
the engine is replaced by a stub whose value for (haplotype h, read r) encodes h and the read's own content, so that the
sample bookkeeping (concatenation into one batch, column ranges, prime / merge / reset) is checked without a GPU."""
import numpy as np
import pytest

from octopus_b200 import HaplotypeLikelihoodArray
from octopus_b200.batch import pack_reads


class StubEngine:
    def __init__(self):
        self.calls = 0

    @staticmethod
    def _read_value(reads, r):
        a, b = int(reads.off[r]), int(reads.off[r + 1])
        return float(int(reads.bases[a:b].astype(np.int64).sum()) * 7 + int(reads.quals[a]) + (b - a))

    def populate(self, config, haps, reads, positions=None, flank_state=None):
        self.calls += 1
        per_read = np.array([self._read_value(reads, r) for r in range(reads.n)])
        return -(np.arange(haps.n)[:, None] * 1e6 + per_read[None, :])

    def populate_templates(self, config, haps, reads, template_off, flank_state=None):
        m = self.populate(config, haps, reads)
        return np.stack([m[:, a:b].sum(axis=1) for a, b in zip(template_off[:-1], template_off[1:])], axis=1)


class Haps:
    n = 3


def _reads(seqs, q):
    return pack_reads([s.encode() for s in seqs], [np.full(len(s), q, np.uint8) for s in seqs])


def test_samples_are_column_ranges_of_one_batch():
    eng = StubEngine()
    s1, s2 = _reads(["ACGT", "GGGTTT"], 30), _reads(["TTTTT"], 20)
    arr = HaplotypeLikelihoodArray(engine=eng).populate({"S1": s1, "S2": s2}, Haps())
    assert eng.calls == 1 and arr.samples() == ["S1", "S2"] and not arr.is_primed()
    assert arr.num_likelihoods("S1") == 2 and arr.num_likelihoods("S2") == 1
    alone1 = StubEngine().populate(None, Haps(), s1)
    alone2 = StubEngine().populate(None, Haps(), s2)
    for h in range(3):
        assert np.array_equal(arr("S1", h), alone1[h]) and np.array_equal(arr("S2", h), alone2[h])
    assert np.array_equal(arr.extract_sample("S2"), alone2)
    with pytest.raises(RuntimeError):
        arr[0]                                   # not primed
    with pytest.raises(KeyError):
        arr("nope", 0)
    arr.prime("S2")
    assert arr.is_primed() and arr.num_likelihoods() == 1 and np.array_equal(arr[2], alone2[2])
    arr.unprime()
    assert not arr.is_primed()


def test_merge_and_reset():
    eng = StubEngine()
    s1, s2, s3 = _reads(["ACGT", "GGGTTT"], 30), _reads(["TTTTT"], 20), _reads(["CA", "AC"], 11)
    arr = HaplotypeLikelihoodArray(engine=eng).populate({"a": s1, "b": s2, "c": s3}, Haps())
    merged = arr.merge_samples()
    assert merged.samples() == ["abc"] and merged.is_primed() and merged.num_likelihoods() == 5
    assert np.array_equal(merged[1], arr.likelihoods[1])
    some = arr.merge_samples(["c", "a"], new_sample="x")
    assert some.samples() == ["x"] and np.array_equal(some[0], np.concatenate([arr("c", 0), arr("a", 0)]))
    before = arr("b", 2).copy()
    arr.reset([0, 2])
    assert arr.likelihoods.shape[0] == 2 and np.array_equal(arr("b", 1), before)
    arr.reset([])
    assert arr.is_empty() and arr.samples() == []


def test_single_block_is_one_primed_sample_and_templates_sum():
    eng = StubEngine()
    s1 = _reads(["ACGT", "GGGTTT", "TT"], 30)
    arr = HaplotypeLikelihoodArray(engine=eng).populate(s1, Haps())
    assert arr.is_primed() and arr.num_likelihoods() == 3 and np.array_equal(arr[1], arr("", 1))
    t = HaplotypeLikelihoodArray(engine=eng).populate_templates({"S": (s1, [0, 2, 3]), "T": (s1, [0, 1, 2, 3])}, Haps())
    assert t.num_likelihoods("S") == 2 and t.num_likelihoods("T") == 3
    m = StubEngine().populate(None, Haps(), s1)
    assert np.array_equal(t("S", 1), [m[1, 0] + m[1, 1], m[1, 2]]) and np.array_equal(t("T", 2), m[2])
