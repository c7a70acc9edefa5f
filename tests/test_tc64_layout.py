"""Host-side check of the index arithmetic of the 64-rows-per-CTA tokeniser kernel (csrc/tc64_layout.cuh, used by
csrc/rq_tc64.cu): staging -> swizzled A image conversion, and the epilogue's view of the M=128 CTA-pair accumulator
layout.  The kernel itself needs a B200 (tests/test_gpu_tc.py); this part of it does not."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_tc64_index_arithmetic(tmp_path):
    exe = tmp_path / "tc64_layout_check"
    src = os.path.join(ROOT, "tests", "host", "tc64_layout_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), src], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert " bad 0 " in res.stdout, res.stdout
