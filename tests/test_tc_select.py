"""Host-side property test of the tensor-core tokeniser's candidate selection core (csrc/tc_select.cuh).

The header compiles for both nvcc and g++; tests/host/tc_select_check.cpp runs the two-stage packed-key top-3 exactly
as the kernel's epilogue does (two 128-column halves, 16-column chunks, merge) against a brute-force model and checks the
safety property the exactness contract rests on: no member of {k : a[k] <= min + margin} is ever lost -- a single
candidate is returned as the argmin, two are returned as (i1, i2), three or more raise `many`.
Score families: wide spread, all-positive / all-negative offsets, clusters at the minimum, exact ties with zero margin,
1e-30 / 1e30 magnitudes and denormals."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_packed_key_selection_never_loses_a_candidate(tmp_path):
    exe = tmp_path / "tc_select_check"
    src = os.path.join(ROOT, "tests", "host", "tc_select_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), src], check=True)
    for mode in ("0", "1"):      # 0: rq_tc_kernel's scan (2 x 128 columns); 1: rq_tc64_kernel's (4 x 64 columns, two scores per insertion)
        res = subprocess.run([str(exe), "200000", mode], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        assert " bad 0 " in res.stdout, res.stdout
