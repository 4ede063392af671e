"""A plain C++ program (g++, no torch) drives the engine through the C ABI and the C++ adapter header."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_cabi.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_cabi")


def _build():
    from octopus_b200 import build
    build.build()
    subprocess.run(["g++", "-std=c++14", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "octopus_b200", "cpp"),
                    SRC, "-o", EXE, "-L", os.path.join(ROOT, "octopus_b200"), "-lphmm_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "octopus_b200")], check=True)


def test_cpp_adapter_compiles_against_the_abi():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_client_matches_oracle(coracle):
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = dict((l.split()[0] + (l.split()[1] if l.startswith("ROW") else ""), l.split()) for l in out.stdout.strip().splitlines())
    assert lines["KAT"][1:] == ["84", "84", "0", "CCCCACGTATATATATATATATGGGGACGT", "CCCCACGT---------------GGGACGT"]
    h0 = "ACGTTGCAAGCTTAGGCTAACGTTAGCATCGATCGGATCTAGCTAGGATCGATACGATCGATCGTAGCTAGCTAGTCGATCGATTTAGCGCGATATCGCGAT"
    h1 = list(h0); h1[50] = "C" if h1[50] == "A" else "A"; h1 = "".join(h1)
    reads = [(h0[30:70], 30, 30), (h1[35:75], 25, 35)]
    n = len(h0)
    for hi, h in enumerate((h0, h1)):
        for ri, (r, q, p) in enumerate(reads):
            st, want, _ = coracle.model_evaluate(16, h, r, np.full(40, q, np.uint8), np.full(n, 40, np.int8), np.full(n, 3, np.int8), h.encode(),
                                                 np.full(n, 50, np.int8), [], p, mapping_quality=60)
            got = float(lines["ROW%d" % hi][2 + ri])
            assert st == 0 and abs(got - want) <= 1e-4 * max(abs(want), 1e-300)
    assert lines["SHORT"][1].startswith("hap=0")
    e2e = dict(kv.split("=") for kv in lines["E2E"][1:])
    assert e2e["same"] == "1" and float(e2e["pinned_ms"]) > 0 and float(e2e["pageable_ms"]) > 0
    assert e2e["gap_open_range"] != "45..45"          # reset() produced repeat-structured penalties (homopolymer / dinucleotide runs)
    print("C++ adapter e2e (64 x 40000 pairs, L=100): pinned %.2f ms, pageable %.2f ms per populate" % (float(e2e["pinned_ms"]), float(e2e["pageable_ms"])))
    assert lines["SAMPLES"][1:] == ["ok", "throws"]          # multi-sample / template container semantics of the C++ adapter


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own hmm::evaluate / hmm::align templates instantiated over GpuPairHMM (tests/cpp/test_dropin.cpp)
# ---------------------------------------------------------------------------------------------------------------------
DROPIN_SRC = os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")
DROPIN_EXE = os.path.join(ROOT, "tests", "cpp", "test_dropin")
REF_SRC = "/root/reference/src"


def _build_dropin():
    """Compiled only where the reference sources exist (never on the GPU box: the binary travels with the snapshot)."""
    if not os.path.isdir(os.path.join(REF_SRC, "core", "models", "pairhmm")):
        return os.path.exists(DROPIN_EXE)
    from octopus_b200 import build
    build.build()
    newest = max(os.path.getmtime(p) for p in (DROPIN_SRC, os.path.join(ROOT, "octopus_b200", "cpp", "phmm_b200.hpp"), os.path.join(ROOT, "include", "phmm_b200.h")))
    if not os.path.exists(DROPIN_EXE) or os.path.getmtime(DROPIN_EXE) < newest:
        subprocess.run(["g++", "-std=c++17", "-O1", "-msse4.1", "-include", "immintrin.h", "-Wno-ignored-attributes",
                        "-I", os.path.join(ROOT, "oracle", "ref_shim"), "-I", REF_SRC, "-I", "/root/reference/lib",
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "octopus_b200", "cpp"),
                        DROPIN_SRC, "-o", DROPIN_EXE, "-L", os.path.join(ROOT, "octopus_b200"), "-lphmm_b200",
                        "-Wl,-rpath," + os.path.join(ROOT, "octopus_b200")], check=True)
    return True


def test_reference_templates_instantiate_over_the_gpu_pair_hmm():
    """hmm::evaluate / hmm::align (pair_hmm.hpp, unmodified) compile with octopus_b200::GpuPairHMM<Band> as their PairHMM."""
    if not os.path.isdir(REF_SRC):
        pytest.skip("reference sources not present (GPU box): the binary was built in the authoring container")
    assert _build_dropin() and os.path.exists(DROPIN_EXE)


@pytest.mark.gpu
def test_reference_evaluate_and_align_over_gpu_pair_hmm_equal_simd_kernel():
    """The reference's evaluate / align code paths (naive shortcut, score-only, traceback + flank discount, CIGAR) driven by the GPU
    kernel give exactly the values they give over the reference's SIMD kernel: 620 seeded cases, bands 8 / 16 / 32."""
    if not _build_dropin():
        pytest.skip("tests/cpp/test_dropin was never built (needs /root/reference at build time)")
    out = subprocess.run([DROPIN_EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "DROPIN ok" in out.stdout, out.stdout + out.stderr
    assert int(out.stdout.split("traceback_cases=")[1].split()[0]) > 100
