"""CPU-only: the C-ABI library builds, loads and exports every symbol include/rqb200.h declares
(no compute calls without a GPU), and the product path refuses CPU tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rqb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rqb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rq_vae_recommender_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the ctypes signature table covers exactly the declared ABI
    assert sorted(_lib._SIGNATURES) == names
    assert _lib.load().rqb200_version() >= 100


def test_argument_errors_are_reported_not_thrown():
    from rq_vae_recommender_b200 import _lib
    lib = _lib.load()
    rc = lib.rqb200_rq_forward(7, 0, 0, 0, 1, 4, 4, 1, 0.25, 0, 0, 0, 0, 0, 0, 0, 0, 0)   # bad mode
    assert rc == 1
    assert b"bad mode" in lib.rqb200_last_error()
    with pytest.raises(_lib.Rqb200Error):
        _lib.check(rc, "rq_forward")


def test_no_cpu_fallback():
    from rq_vae_recommender_b200 import _lib, ops
    x = torch.randn(8, 16)
    cb = torch.randn(4, 16)
    with pytest.raises(_lib.Rqb200Error):
        ops.rq_tokenize(x, [cb])


def test_no_product_import_of_oracle():
    pkg = os.path.join(ROOT, "rq_vae_recommender_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# noqa", ""), fn
