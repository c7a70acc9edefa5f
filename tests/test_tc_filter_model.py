"""CPU check of the DETERMINISTIC filter margin of the tensor-core tokeniser (tests/tc_filter_model.py restates the kernel's
formulas).  Properties: (1) the observed fp16 score error never exceeds eps -- on gaussian workloads AND on inputs built to
make the rounding errors coherent (the family that defeats a z-sigma margin, incl. the round-1 judge's counterexample);
(2) the float64 argmin and the oracle's fp32 argmin are never filtered out; (3) the candidate sets stay small on the
workloads the GPU tests and bench.py use (the re-rank is the rare path)."""
import numpy as np
import pytest

import inputs as I
import tc_filter_model as M
from oracle import rq_oracle as O


def _run(x, cbs):
    ids = O.rq_tokenize(x, cbs)
    lv = M.filter_levels(x, cbs, ids)
    worst, frac = 0.0, []
    for l, r in enumerate(lv):
        rows = np.arange(len(x))
        assert r["cand"][rows, ids[:, l]].all(), f"level {l}: the oracle's fp32 argmin was filtered out"
        true = M.true_half_distances(x, cbs, ids, l)
        fin = np.isfinite(r["eps"]) & np.isfinite(true).all(1)
        assert r["cand"][rows[fin], true[fin].argmin(1)].all(), f"level {l}: the float64 argmin was filtered out"
        with np.errstate(invalid="ignore", over="ignore"):
            err = np.abs(r["h"].astype(np.float64) - true)
        if fin.any():
            worst = max(worst, float((err[fin].max(1) / r["eps"][fin]).max()))
        assert r["cand"][~fin].all(), "rows with non-finite statistics must keep every code"
        frac.append(float((r["cand"].sum(1) > 1).mean()))
    return worst, frac


@pytest.mark.parametrize("n,D,L,seed", [(4096, 768, 3, 1234), (2048, 256, 4, 7), (2048, 64, 3, 11)])
def test_filter_keeps_the_exact_argmin(n, D, L, seed):
    x, cbs = I.rq_problem(n, D, 256, L, seed=seed)
    worst, frac = _run(x, cbs)
    assert worst <= 1.0, worst           # a bound, not a statistic: observed error / eps can never exceed 1
    assert max(frac) < 0.12, frac        # rows needing the exact re-rank per level


@pytest.mark.parametrize("kind", M.ADVERSARIAL_KINDS)
@pytest.mark.parametrize("D,L", [(768, 1), (768, 3), (128, 2)])
def test_filter_bound_holds_on_coherent_rounding(kind, D, L):
    x, cbs = M.adversarial_problem(kind, D=D, L=L, n=96)
    worst, _ = _run(x, cbs)
    assert worst <= 1.0, (kind, worst)


def test_judge_counterexample_is_flagged():
    """fp32/fp64 say code 10, a 4.5-sigma margin keeps only code 200.  The bound keeps both."""
    x, cbs = M.adversarial_problem("judge_r1", D=768, L=1, n=4)
    ids = O.rq_tokenize(x, cbs)
    assert (ids[:, 0] == 10).all()
    r = M.filter_levels(x, cbs, ids)[0]
    assert r["cand"][:, 10].all() and r["cand"][:, 200].all()
    assert (r["h"][:, 200] < r["h"][:, 10]).all()      # the fp16 scores alone would pick the wrong code


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
