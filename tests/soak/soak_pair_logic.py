"""Long-running CPU soak (test infrastructure, not collected by pytest): the populate kernels' per-pair logic (candidate slots, shortcut, DP, flank discount,
mapping-quality mixing; the engine's device functions compiled for the CPU) against the plain-C oracle on adversarial inputs.

usage: python tests/soak/soak_pair_logic.py [seed] [cases]
"""
import os, sys, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle.oracle import COracle
co = COracle()
import os
lib = C.CDLL(os.path.join(ROOT, 'tests', 'cpu_emul', 'libphmm_emul.so'))
vp = C.c_void_p
lib.emul_pair_evaluate.argtypes = [C.c_int, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
P = lambda a: a.ctypes.data
seed = int(sys.argv[1]); N = int(sys.argv[2])
rng = np.random.default_rng(seed)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
bad = 0
for it in range(N):
    band = int(rng.choice([8, 16, 32]))
    L = int(rng.integers(6, 120))
    Lh = int(rng.integers(L + 2 * band - 4, L + 2 * band + 150))
    style = rng.integers(0, 3)
    if style == 0: hap = acgt[rng.integers(0, 4, Lh)].copy()
    elif style == 1:
        unit = acgt[rng.integers(0, 4, int(rng.integers(1, 5)))]; hap = np.tile(unit, Lh // len(unit) + 1)[:Lh].copy()
        for _ in range(int(rng.integers(0, 5))): hap[rng.integers(0, Lh)] = acgt[rng.integers(0, 4)]
    else:
        hap = acgt[rng.integers(0, 4, Lh)].copy(); hap[rng.random(Lh) < 0.05] = ord("N")
    mask = np.roll(hap, 1) if rng.random() < 0.5 else np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, Lh)].copy()
    ext = rng.random() < 0.3
    prior = (rng.choice([0, 1, 125], Lh) if ext else rng.integers(0, 126, Lh)).astype(np.int8)
    go = (rng.choice([0, 1, 45], Lh) if ext else rng.integers(0, 46, Lh)).astype(np.int8)
    ge = (rng.choice([0, 1, 10], Lh) if ext else rng.integers(0, 11, Lh)).astype(np.int8)
    p0 = int(rng.integers(0, max(1, Lh - L + 1)))
    read = hap[p0:p0 + L].copy()
    if len(read) < L: read = np.concatenate([read, acgt[rng.integers(0, 4, L - len(read))]])
    read[read == ord("N")] = ord("A")
    for _ in range(int(rng.choice([0, 0, 1, 1, 2, 3, 5]))): read[rng.integers(0, L)] = acgt[rng.integers(0, 4)]
    if rng.random() < 0.2:
        i = int(rng.integers(1, L - 1)); read = np.concatenate([read[:i], read[i + 1:], acgt[rng.integers(0, 4, 1)]])
    q = (rng.integers(0, 94, L) if rng.random() < 0.7 else rng.choice([0, 1, 2, 93], L)).astype(np.uint8)
    npos = int(rng.integers(0, 5))
    pos = np.array([min(max(0, p0 + int(rng.integers(-20, 21))), Lh) for _ in range(npos)], dtype=np.int32)
    if npos and rng.random() < 0.5: pos[rng.integers(0, npos)] = p0
    orig = max(0, p0 + int(rng.choice([0, 0, 0, -3, 4, -40, 40])))
    uf = int(rng.random() < 0.6)
    lhs, rhs = int(rng.integers(0, Lh // 2)), int(rng.integers(0, Lh // 2))
    mq = int(rng.choice([0, 10, 29, 60, 255]))
    usemq, trig, dpo = int(rng.random() < 0.8), int(rng.choice([-1, 40, 200])), int(rng.random() < 0.3)
    cap = int(rng.choice([120, 50]))
    out, ex = C.c_double(0), C.c_int(0)
    rc = lib.emul_pair_evaluate(band, P(hap), Lh, P(mask), P(prior), P(go), P(ge), P(read), P(q), L, uf, lhs, rhs, P(pos) if npos else None, npos, orig, usemq, mq, cap, trig, dpo, 2, C.byref(out), C.byref(ex))
    st, val, e = co.model_evaluate(band, hap.tobytes(), read.tobytes(), q, go, ge, mask.tobytes(), prior, pos.astype(np.int64), orig, mapping_quality=mq, flanks=(lhs, rhs) if uf else None, use_mapping_quality=bool(usemq), mapq_cap=cap, mapq_cap_trigger=trig, dp_only=bool(dpo))
    ok = rc == st and ((rc == 0 and out.value == val) or (rc != 0 and ex.value == e))
    if not ok:
        bad += 1
        if bad <= 5: print('MISMATCH', seed, it, band, L, Lh, (rc, out.value, ex.value), (st, val, e), 'uf', uf, lhs, rhs, 'dpo', dpo, flush=True)
print('seed', seed, 'cases', N, 'bad', bad, flush=True)
