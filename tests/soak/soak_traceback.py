"""Long-running CPU soak (test infrastructure, not collected by pytest): the engine's traceback path (generic_align<true>: score, first_pos, flank replay) compiled
for the CPU against the oracle, and the oracle's traceback (both alignment strings) against the reference's own SIMD kernel.

usage: python tests/soak/soak_traceback.py [seed] [cases]
"""
import os, sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'soak'))
import soak_emul as S
emul = S.emul
vp = C.c_void_p
emul.emul_generic.argtypes = [C.c_int, C.c_int, C.c_int] + [vp] * 7 + [C.c_int, C.c_int, C.c_int, vp, vp, vp]
co = S.COracle(); ref = S.RefKernel(S.available_ref_isas()[0])
seed = int(sys.argv[1]); N = int(sys.argv[2])
rng = np.random.default_rng(seed)
bad = 0
P = S.P
for it in range(N):
    band = int(rng.choice([8, 16, 32, 64]))
    L = int(rng.integers(1, 150))
    nuc = int(rng.choice([0, 1, 2, 5]))
    c = S.case(rng, band, L)
    if rng.random() < 0.2 and L > 2: c["read"][rng.integers(0, L)] = ord("N")
    W = len(c["truth"])
    lhs, rhs = int(rng.integers(0, W + 1)) // 2, int(rng.integers(0, W + 1)) // 2
    q8 = c["quals"].astype(np.int8)
    fp, fs, ms = C.c_int(0), C.c_int(0), C.c_int(0)
    args = (L, P(c["read"]), P(q8), P(c["truth"]), P(c["snv_mask"]), P(c["snv_prior"]), P(c["gap_open"]), P(c["gap_extend"]), nuc)
    st = emul.emul_generic(band, 1, *args, lhs, rhs, C.byref(fp), C.byref(fs), C.byref(ms))
    t, r, m = c["truth"].tobytes(), c["read"].tobytes(), c["snv_mask"].tobytes()
    es, efp, a1, a2 = co.align_tb(band, t, r, q8, c["gap_open"], c["gap_extend"], nuc, m, c["snv_prior"])
    efs, ems = co.flank_score(W, lhs, rhs, r, q8, m, c["snv_prior"], c["gap_open"], c["gap_extend"], nuc, efp, a1, a2)
    if (st, fp.value, fs.value, ms.value) != (es, efp, efs, ems):
        bad += 1; print('EMUL-vs-ORACLE', seed, it, band, L, (st, fp.value, fs.value, ms.value), (es, efp, efs, ems), flush=True)
    if es < 7000:
        rs, rfp, ra1, ra2 = ref.align_tb(band, t, r, q8, c["gap_open"], c["gap_extend"], nuc, m, c["snv_prior"], bits=32)
        if (rs, rfp, ra1, ra2) != (es, efp, a1, a2):
            bad += 1; print('ORACLE-vs-REF tb', seed, it, band, L, (es, efp, a1, a2), (rs, rfp, ra1, ra2), flush=True)
        else:
            rfs, rms = ref.flank_score(band, W, lhs, rhs, r, q8, m, c["snv_prior"], c["gap_open"], c["gap_extend"], nuc, rfp, ra1, ra2, bits=32)
            if (rfs, rms) != (efs, ems):
                bad += 1; print('ORACLE-vs-REF flank', seed, it, band, L, (efs, ems), (rfs, rms), flush=True)
    if bad > 10: break
print('seed', seed, 'cases', it + 1, 'bad', bad, flush=True)
