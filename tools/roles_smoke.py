#!/usr/bin/env python
"""Tiny band-32 / band-64 populate on the role-warps kernel (k_populate_roles) against the lane kernel's values; small enough to run
under `compute-sanitizer --tool racecheck` (shared-memory hand-overs between the warps of a group)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import random_region          # noqa: E402
from octopus_b200 import HaplotypeLikelihoodModel, PairHMMEngine   # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(3)
eng = PairHMMEngine(0)
for band in (32, 64):
    haps, reads = random_region(rng, band, n_haps=20, n_reads=n_reads, hap_len=2 * band + 200, read_len_choices=[40, 76, 100], read_n_rate=0.0)
    cfg = HaplotypeLikelihoodModel.Config(max_indel_error=band, disable_naive_shortcut=True, map_positions=False)
    got = eng.populate(cfg, haps, reads)
    print("band", band, "reads", reads.n, "haps", haps.n, "checksum", float(np.nan_to_num(got, neginf=-1e9).sum()), "launches", eng.launch_count())
