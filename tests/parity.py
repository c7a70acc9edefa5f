"""Parity protocol helpers shared by the CPU and GPU tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


# ---- the parity protocol of SURVEY 8(c): ids exact except on float64-classified near ties -------------
TAU = 1e-5
TAU_ABS = 1e-6      # ~16 fp32 ulps of the operands xx + cc: below that the reference's own fp32 distances are summation-order noise


def assert_ids_match(ids, ref_ids, res, codebooks, what=""):
    """ids/ref_ids [B,L].  Rows may differ from the fp32 reference only where the float64 oracle says the
    top-2 gap is a near tie at the first differing level -- relative gap (d2 - d1) / d1 <= TAU, or, for rows that all but
    coincide with a code (d1 is then a cancellation residue), (d2 - d1) <= TAU_ABS (xx + cc) -- and then only by
    picking the runner-up.  Returns the number of such near-tie rows."""
    from oracle import rq_oracle as O
    ids = np.asarray(ids).astype(np.int64).reshape(len(ids), -1)
    ref_ids = np.asarray(ref_ids).astype(np.int64).reshape(len(ref_ids), -1)
    bad = np.nonzero((ids != ref_ids).any(axis=1))[0]
    if len(bad) == 0:
        return 0
    cbs64 = [np.asarray(c, np.float64) for c in codebooks]
    r64 = np.asarray(res, np.float64)[bad]
    # both chains are identical up to the first differing level l, so one float64 pass following impl is enough
    i64, second, gap, absgap = O.top2_gap(r64, cbs64, ids[bad], return_abs=True)
    for r in range(len(bad)):
        l = int(np.nonzero(ids[bad[r]] != ref_ids[bad[r]])[0][0])
        pair = {int(i64[r, l]), int(second[r, l])}
        assert (gap[r, l] <= TAU or absgap[r, l] <= TAU_ABS) and int(ids[bad[r], l]) in pair and int(ref_ids[bad[r], l]) in pair, (
            f"{what}: row {bad[r]} level {l}: impl {ids[bad[r]]} vs ref {ref_ids[bad[r]]}; "
            f"fp64 says {i64[r]} / runner-up {second[r]} gap {gap[r, l]:.3e} (abs {absgap[r, l]:.3e})")
    return len(bad)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
