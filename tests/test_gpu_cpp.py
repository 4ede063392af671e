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
    assert lines["SAMPLES"][1:] == ["ok", "throws"]          # multi-sample / template container semantics of the C++ adapter
