import sys, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from test_emul import _fb_case, P, vp
from oracle.oracle import COracle
from octopus_b200.build import build_cpu_emulation
emul = C.CDLL(build_cpu_emulation())
emul.emul_dp_flank_fb.argtypes = [C.c_int, C.c_int] + [vp] * 14 + [C.c_int] * 5 + [vp, vp]
emul.emul_force_form.argtypes = [C.c_int]
co = COracle()
seed = int(sys.argv[1]); n = int(sys.argv[2])
rng = np.random.default_rng(seed)
used = tie = quirk = 0
for it in range(n):
    band = int(rng.choice([8, 16, 32], p=[0.4, 0.45, 0.15]))
    L = int(rng.integers(2 * band, 2 * band + 200))
    nuc = int(rng.integers(0, 5))
    ordered = it % 2 == 0
    a, la, ra = _fb_case(rng, band, L, it, qmax=60 if it % 7 == 0 else 41, ordered=ordered)
    b, lb, rb = _fb_case(rng, band, L, it + 1, ordered=ordered)
    if it % 3 == 0: lb, rb = la, ra
    o0, o1 = (C.c_int * 4)(), (C.c_int * 4)()
    emul.emul_force_form(0 if it % 4 == 0 else -1)
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
        if o[3]: tie += 1; continue
        if (o[1], o[2]) != (efs, ems):
            hasn = bool(((c["truth"] == ord("N")) & (c["snv_prior"] < 2)).any())
            assert hasn and o[2] == ems and o[1] <= efs, (it, band, L, lhs, rhs, tuple(o), (es, efs, ems))
            quirk += 1
print("seed", seed, "used", used, "tie", tie, "quirk", quirk)
