"""Long-running CPU soak (not part of the test suite): dp_flank_fb — the packed forward / backward flank core the kernels k_flank_fwd /
k_flank_bwd run — compiled for the CPU with the DPX instructions emulated (tests/cpu_emul), against the oracle's traceback +
calculate_flank_score. Two alignments per call with independent flank geometries; random inputs (tests/helpers.py) or, with a third
argument, the adversarial generator of soak_emul.py (low-complexity and N-rich windows, penalties at the extremes of the error-model
ranges and 0, qualities 0 .. the 16-bit bound).
A case counts as checked when the core reports no tie (ties go to the labelled DP in the kernels) and the replay quirk does not apply
(truth 'N' matchable below 2: flank_replay_may_differ sends those to the traceback kernel).

usage: python tests/soak/soak_flank_fb.py seed cases [adversarial]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "soak"))
from test_emul import _fb_case, P, vp                       # noqa: E402
from oracle.oracle import COracle                           # noqa: E402
from octopus_b200.build import build_cpu_emulation          # noqa: E402

emul = C.CDLL(build_cpu_emulation())
emul.emul_dp_flank_fb.argtypes = [C.c_int, C.c_int] + [vp] * 14 + [C.c_int] * 5 + [vp, vp]
emul.emul_force_form.argtypes = [C.c_int]
co = COracle()
seed, n = int(sys.argv[1]), int(sys.argv[2])
adversarial = len(sys.argv) > 3
if adversarial:
    from soak_emul import case as adv_case                  # noqa: E402
rng = np.random.default_rng(seed)


def geometry(W, band, it):
    mode = it % 5
    if mode == 0:
        lhs, rhs = int(rng.integers(0, W // 2 + 1)), int(rng.integers(0, W // 2 + 1))
    elif mode == 1:
        lhs, rhs = int(rng.integers(1, W)), 0
    elif mode == 2:
        lhs, rhs = 0, int(rng.integers(1, W))
    elif mode == 3:
        lhs, rhs = int(rng.integers(0, W + 1)), int(rng.integers(0, W + 1))
    else:
        lhs, rhs = int(rng.integers(0, 2 * band + 2)), int(rng.integers(0, 2 * band + 2))
    if W - rhs <= lhs or (lhs == 0 and rhs == 0):
        lhs, rhs = max(1, W // 4), 0
    return lhs, rhs


used = tie = quirk = 0
for it in range(n):
    band = int(rng.choice([8, 16, 32], p=[0.4, 0.45, 0.15]))
    L = int(rng.integers(2 * band, 2 * band + 200))
    nuc = int(rng.integers(0, 5))
    if adversarial:
        a, b = adv_case(rng, band, L), adv_case(rng, band, L)
        (la, ra), (lb, rb) = geometry(len(a["truth"]), band, it), geometry(len(b["truth"]), band, it + 1)
        emul.emul_force_form(-1)
    else:
        ordered = it % 2 == 0
        a, la, ra = _fb_case(rng, band, L, it, qmax=60 if it % 7 == 0 else 41, ordered=ordered)
        b, lb, rb = _fb_case(rng, band, L, it + 1, ordered=ordered)
        emul.emul_force_form(0 if it % 4 == 0 else -1)
    if it % 3 == 0:
        lb, rb = la, ra
    o0, o1 = (C.c_int * 4)(), (C.c_int * 4)()
    rc = emul.emul_dp_flank_fb(band, L, P(a["read"]), P(a["quals"]), P(b["read"]), P(b["quals"]),
                               P(a["truth"]), P(a["snv_mask"]), P(a["snv_prior"]), P(a["gap_open"]), P(a["gap_extend"]),
                               P(b["truth"]), P(b["snv_mask"]), P(b["snv_prior"]), P(b["gap_open"]), P(b["gap_extend"]),
                               nuc, la, ra, lb, rb, o0, o1)
    assert rc == 0
    for c, lhs, rhs, o in ((a, la, ra, o0), (b, lb, rb, o1)):
        W = len(c["truth"]); q8 = c["quals"].astype(np.int8)
        t, r, m = c["truth"].tobytes(), c["read"].tobytes(), c["snv_mask"].tobytes()
        es, efp, a1, a2 = co.align_tb(band, t, r, q8, c["gap_open"], c["gap_extend"], nuc, m, c["snv_prior"])
        efs, ems = co.flank_score(W, lhs, rhs, r, q8, m, c["snv_prior"], c["gap_open"], c["gap_extend"], nuc, efp, a1, a2)
        used += 1
        assert o[0] == es, (it, band, L, lhs, rhs, o[0], es)
        if o[3]:
            tie += 1
            continue
        if (o[1], o[2]) != (efs, ems):
            may_quirk = bool((c["truth"] == ord("N")).any()) and (bool((c["snv_prior"] < 2).any()) or bool((c["quals"] < 2).any()))
            assert may_quirk and o[2] == ems and o[1] <= efs, (it, band, L, lhs, rhs, tuple(o), (es, efs, ems))
            quirk += 1
emul.emul_force_form(-1)
print("seed", seed, "adversarial" if adversarial else "random", "alignments", used, "ties", tie, "replay-quirk cases", quirk, "mismatches 0")
