"""CPU check of the statistical filter margin of the tensor-core tokeniser (tests/tc_filter_model.py restates the kernels'
formulas).  Property: on the workloads the GPU tests and bench.py use, the oracle's fp32 argmin is never filtered out, the
observed fp16 error stays well inside eps, and the candidate sets stay small (the re-rank is the rare path)."""
import numpy as np
import pytest

import inputs as I
import tc_filter_model as M
from oracle import rq_oracle as O


def _run(x, cbs, gram16=False):
    ids = O.rq_tokenize(x, cbs)
    lv = M.filter_levels(x, cbs, ids, gram16=gram16)
    worst, frac = 0.0, []
    for l, r in enumerate(lv):
        assert r["cand"][np.arange(len(x)), ids[:, l]].all(), f"level {l}: the exact argmin was filtered out"
        err = np.abs(r["h"].astype(np.float64) - M.true_half_distances(x, cbs, ids, l))
        worst = max(worst, float((err.max(1) / r["eps"]).max()))
        frac.append(float((r["cand"].sum(1) > 1).mean()))
    return worst, frac


@pytest.mark.parametrize("n,D,L,seed", [(4096, 768, 3, 1234), (2048, 256, 4, 7), (2048, 64, 3, 11)])
def test_filter_keeps_the_exact_argmin(n, D, L, seed):
    x, cbs = I.rq_problem(n, D, 256, L, seed=seed)
    worst, frac = _run(x, cbs)
    assert worst < 0.6, worst            # observed error / eps: the margin has room (z = 4.5 statistical term dominates)
    assert max(frac) < 0.08, frac        # rows needing the exact re-rank per level


def test_filter_scaled_rows_and_overflow():
    x, cbs = I.rq_problem(2048, 768, 256, 3, seed=99)
    x = x[:512].copy()
    x[0:64] *= 1e-3
    x[64:128] *= 37.0
    x[128:132] *= 1e6                    # fp16 overflow: every code must become a candidate
    ids = O.rq_tokenize(x, cbs)
    lv = M.filter_levels(x, cbs, ids)
    for l, r in enumerate(lv):
        assert r["cand"][np.arange(len(x)), ids[:, l]].all()
    assert lv[0]["cand"][128:132].all()
