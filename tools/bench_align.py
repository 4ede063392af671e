#!/usr/bin/env python
"""Throughput of phmm_align_reads (HaplotypeLikelihoodModel::align: traceback + CIGAR per pair) on a C2-shaped batch:
one (read, haplotype) pair per read, the read's original position as the only candidate. Prints one JSON line.
PHMM_NO_FAST_ALIGN=1 in the environment measures the generic (round-1) traceback kernel instead."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from octopus_b200 import HaplotypeLikelihoodModel, PairHMMEngine, synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    haps, reads, band = synth.make_batch("C2", n_reads=n)
    rng = np.random.default_rng(5)
    pairs = np.stack([np.arange(n, dtype=np.int32), rng.integers(0, haps.n, n).astype(np.int32)], axis=1)
    cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band)
    eng = PairHMMEngine(0)
    eng.align_reads(cfg, haps, reads, pairs[:1000])
    t0 = time.perf_counter()
    mp, lk, cig, st = eng.align_reads(cfg, haps, reads, pairs)
    dt = time.perf_counter() - t0
    kernel_ms = eng.last_dp_kernel_ms()
    cells = int((2 * (np.diff(reads.off) + band) * band).sum())
    exact = sum(1 for c in cig if c.endswith("=") and c[:-1].isdigit())
    print(json.dumps({"what": "phmm_align_reads", "pairs": n, "band": band, "kernels_ms": kernel_ms, "call_ms": 1e3 * dt,
                      "gcups_kernels": cells / (kernel_ms / 1e3) / 1e9, "exact_match_pairs": exact, "status_ok": int((st == 0).sum()),
                      "generic_kernel_forced": bool(os.environ.get("PHMM_NO_FAST_ALIGN"))}))


if __name__ == "__main__":
    main()
