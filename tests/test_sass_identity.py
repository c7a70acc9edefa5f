"""The opt-in tokeniser variants (extra template parameters / `if constexpr` branches of rq_tc_kernel) must not
change the code that runs by default: the SASS of every shipped instantiation has to be byte-identical to the last commit whose
default path was validated on hardware.  When the default kernel is changed on purpose (and re-validated on a B200), update REF."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "0ac61f5"      # end of the hardware-validated part of round 1 (profiles/r1_bench_final.json was measured with this code)


def _have_ref():
    if shutil.which("git") is None or shutil.which("nvcc") is None or shutil.which("cuobjdump") is None:
        return False
    return subprocess.run(["git", "-C", ROOT, "cat-file", "-e", REF + "^{commit}"], capture_output=True).returncode == 0


@pytest.mark.skipif(not _have_ref(), reason="needs git history with the reference commit, nvcc and cuobjdump")
def test_shipped_instantiations_are_unchanged():
    res = subprocess.run(["bash", os.path.join(ROOT, "tools", "sass_identity.sh"), REF], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = [l for l in res.stdout.splitlines() if "->" in l]
    assert len(lines) == 8 and all(l.endswith(" 0 differing lines") for l in lines), res.stdout
