"""numpy model of the tensor-core tokeniser's candidate FILTER (csrc/rq_tc.cu / rq_tc64.cu) -- test infrastructure.

The kernels' exactness argument has two halves: the exact fp32 re-rank (same arithmetic as the CUDA-core kernel, tested
against the oracle on the GPU) and the claim that the fp16 tensor-core scores never drop the true argmin from the candidate
set {k : h[k] <= min h + 2 eps_b}.  The second half is statistical (DESIGN.md 5.2 "filter error bound") and this model
restates it on the CPU with the kernel's own formulas, so the margin can be checked -- and changed -- without a GPU:

  x~ = fp16(x), c~ = fp16(c * 2^s)              (tc_prep_blob_kernel, converter)
  S  = x~ . c~                                  (tcgen05.mma, fp32 accumulation: modelled as an exact sum rounded once)
  h  = T[k] - S / 2^s,  T = cc/2 + sum_j G_jl[id_j]   (tc_prep_fold_kernel; epilogue score16)
  eps_b = Z u sqrt(2/3) sqrt(max|x| ||x||_2 c4max) + 2^-25 (c1max + sqrt(D) ||x|| / 2^s) + 2^-17 ||x|| c2max + gerr
  candidates = {k : h[k] <= tcs_threshold(min h, 2 eps_b)}     (csrc/tc_select.cuh)

`gram16=True` models the planned fp16 Gram tables (power-of-two scaled, cc/2 kept in fp32, 2^-11 max|G| added to gerr).
"""
import numpy as np

TC_Z = 4.5
U16 = 2.0 ** -11


def _bf16_up(v):
    b = np.asarray(v, np.float32).view(np.uint32).copy()
    fin = (b & 0x7F800000) != 0x7F800000
    b[fin & ((b & 0xFFFF) != 0)] += 0x10000
    return (b & 0xFFFF0000).view(np.float32)


def prepare(cbs):
    """Per-level constants of tc_prep_stats_kernel / tc_prep_consts_kernel and the fp16 codebook images."""
    lv = []
    for l, c in enumerate(cbs):
        c = np.asarray(c, np.float32)
        amax = float(np.abs(c).max())
        sc = 1.0
        if amax > 0 and np.isfinite(amax):
            e = int(np.clip(np.frexp(amax)[1], -60, 60))
            sc = float(np.ldexp(1.0, -e))
        c64 = c.astype(np.float64)
        lv.append(dict(sc=sc, c4max=float(np.sqrt((c64 ** 4).sum(1)).max()), c1max=float(np.abs(c64).sum(1).max()),
                       c2max=float(np.sqrt((c64 ** 2).sum(1)).max()), cc=(c * c).sum(1, dtype=np.float32),
                       img=(c * np.float32(sc)).astype(np.float16)))
    for l in range(len(cbs)):
        lv[l]["gerr"] = 3.8e-6 * sum(lv[j]["c2max"] for j in range(l)) * lv[l]["c2max"]
    return lv


def gram_tables(cbs, gram16=False):
    """G[(j, l)] = C_j C_l^T in fp32; with gram16 the values pass through a power-of-two scaled fp16 and the rounding
    bound 2^-11 max|G| is returned alongside."""
    out = {}
    for l in range(1, len(cbs)):
        for j in range(l):
            g = np.asarray(cbs[j], np.float32) @ np.asarray(cbs[l], np.float32).T
            err = 0.0
            if gram16:
                gmax = float(np.abs(g).max())
                s = float(np.ldexp(1.0, 14 - np.frexp(gmax)[1])) if gmax > 0 else 1.0      # max|G| s in [2^13, 2^14)
                g = ((g * np.float32(s)).astype(np.float16).astype(np.float32) / np.float32(s)).astype(np.float32)
                err = U16 * gmax
            out[(j, l)] = (g, err)
    return out


def filter_levels(x, cbs, ids, z=TC_Z, gram16=False):
    """ids: the exact chain's ids [B, L] (the kernel feeds the FINAL ids of earlier levels into the Gram correction).
    Returns per level: candidate mask [B, K], eps [B], approximate half-distances h [B, K]."""
    x = np.asarray(x, np.float32)
    B, D = x.shape
    lv = prepare(cbs)
    grams = gram_tables(cbs, gram16)
    xh = x.astype(np.float16)
    xmax = _bf16_up(np.abs(x).max(1))
    x2s = _bf16_up((x.astype(np.float64) ** 2).sum(1).astype(np.float32))
    overflow = ~(xmax < 65504.0)
    x4s = np.where(overflow, np.inf, xmax.astype(np.float64) ** 2 * x2s)
    x2n = np.sqrt(x2s.astype(np.float64))
    out = []
    for l, c in enumerate(cbs):
        k = lv[l]
        S = (xh.astype(np.float64) @ k["img"].astype(np.float64).T).astype(np.float32)
        T = np.broadcast_to(0.5 * k["cc"], (B, len(k["cc"]))).astype(np.float32).copy()
        gerr = k["gerr"]
        for j in range(l):
            g, e = grams[(j, l)]
            T += g[ids[:, j]]
            gerr += e
        h = (T - S * np.float32(1.0 / k["sc"])).astype(np.float32)
        sig = U16 * 0.81649658 * np.sqrt(np.sqrt(x4s) * k["c4max"])
        flo = 2.98023224e-8 * (k["c1max"] + np.sqrt(D) * x2n / k["sc"])
        acc = 7.62939453e-6 * x2n * k["c2max"]
        eps = z * sig + flo + acc + gerr
        m1 = h.min(1)
        thr = m1 + (2.0 * eps * 1.00001 + 4.5e-6 * np.abs(m1) + 1e-43)
        cand = ~(h > thr[:, None])
        out.append(dict(cand=cand, eps=eps, h=h))
    return out


def true_half_distances(x, cbs, ids, level):
    """float64 half-distance cc/2 - res.c of `level` along the chain given by ids."""
    res = np.asarray(x, np.float64).copy()
    for j in range(level):
        res -= np.asarray(cbs[j], np.float64)[ids[:, j]]
    c = np.asarray(cbs[level], np.float64)
    return 0.5 * (c * c).sum(1)[None, :] - res @ c.T
