"""CPU check of the device math header (oscen_b200/csrc/om_math.h) against the platform libm.

om_math.h is compiled here as plain C++ (the same source nvcc compiles for the device; every
operation in it is an IEEE add/mul/fma/div/convert that the GPU evaluates identically), and swept
against glibc on a strided subset of all f32 bit patterns.  The exhaustive sweep (stride 1) is
recorded in DESIGN.md."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def results(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("om") / "check_om_math")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "csrc", "check_om_math.cpp"), "-lpthread", "-lm"])
    out = subprocess.check_output([exe, "211", "4"], text=True, timeout=300)
    res = {}
    for line in out.splitlines():
        m = re.match(r"(.+?) n=(\d+) mismatch=(\d+)(?: max_ulp=(\d+))?", line)
        assert m, line
        res[m.group(1).split(" ")[0]] = (int(m.group(2)), int(m.group(3)), int(m.group(4) or 0))
    return res


@pytest.mark.parametrize("name", ["expf[|x|<=104]", "sinf[|x|<120]", "cosf[|x|<120]",
                                  "sinf[120<=|x|<inf]", "cosf[120<=|x|<inf]", "tanhf[|x|<=12]", "tanhf_u[|x|<=32]",
                                  "sinf_u[all]", "cosf_u[all]", "sinf_r[|x|<120]",
                                  "expf[adsr"])
def test_bit_exact_against_glibc(results, name):
    n, bad, _ = results[name]
    assert n > 1_000_000 and bad == 0, (name, n, bad)


def test_tanf_within_tolerance(results):
    n, bad, max_ulp = results["tanf[|x|<=2]"]
    assert bad / n < 1e-5 and max_ulp <= 2


def test_tanf_mid_range_core_equals_the_general_routine(results):
    """om_k_tanf_mid (the branch-free core the segmented FM kernel inlines) is bit-identical to om_tanf on its whole range"""
    n, bad, _ = results["tanf_mid[2^-13<=x<0.6744]"]
    assert n > 400_000 and bad == 0, (n, bad)
