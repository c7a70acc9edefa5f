"""Seeded synthetic inputs shared by the golden generator (build container) and the tests (anywhere).

Everything comes from numpy's frozen legacy ``RandomState`` streams and IEEE-exact elementwise
arithmetic, so the GPU box regenerates bit-identical inputs; every fixture stores a sha1 of the
regenerated arrays and the tests assert it before comparing outputs.
"""
import hashlib

import numpy as np


def sha(*arrays) -> str:
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def randn(seed: int, *shape) -> np.ndarray:
    return np.random.RandomState(seed).randn(*shape).astype(np.float32)


def rand(seed: int, *shape) -> np.ndarray:
    return np.random.RandomState(seed).rand(*shape).astype(np.float32)


def unit_rows(seed: int, n: int, d: int) -> np.ndarray:
    """Unit-norm rows (sentence-T5-like item features), normalised in float64 then cast."""
    x = np.random.RandomState(seed).randn(n, d)
    x /= np.sqrt((x * x).sum(axis=1, keepdims=True))
    return x.astype(np.float32)


def _argmin64(res: np.ndarray, cb: np.ndarray) -> np.ndarray:
    r, c = res.astype(np.float64), cb.astype(np.float64)
    out = np.empty(res.shape[0], np.int64)
    for s in range(0, res.shape[0], 8192):
        d = (c * c).sum(1)[None, :] - 2.0 * (r[s:s + 8192] @ c.T)
        out[s:s + 8192] = np.argmin(d, axis=1)
    return out


def rq_problem(n: int, d: int = 768, k: int = 256, L: int = 3, seed: int = 1234,
               noise: float = 0.5, x: np.ndarray = None):
    """Items + L 'live' codebooks: level-l codes are residual rows of that level plus noise/sqrt(d)
    gaussian jitter, so (like a k-means-initialised model) every code attracts rows."""
    if x is None:
        x = unit_rows(seed, n, d)
    rs = np.random.RandomState(seed + 1)
    cbs = []
    res = x.copy()
    for _ in range(L):
        idx = rs.choice(n, k, replace=False)
        jitter = (rs.randn(k, d) * (noise / np.sqrt(d))).astype(np.float32)
        cb = (res[idx] + jitter).astype(np.float32)
        cbs.append(cb)
        res = res - cb[_argmin64(res, cb)]
    return x, cbs


def mlp_weights(seed: int, dims):
    """nn.Linear-layout [out,in] weights, U(-1/sqrt(in), 1/sqrt(in)) like torch's default init."""
    rs = np.random.RandomState(seed)
    ws = []
    for i, o in zip(dims[:-1], dims[1:]):
        b = 1.0 / np.sqrt(i)
        ws.append(((rs.rand(o, i) * 2 - 1) * b).astype(np.float32))
    return ws
