"""torch.Tensor <-> librqb200 C-ABI marshalling and the autograd Functions built on it.

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every FLOP of the hot path
runs in the hand-written kernels of csrc/.  CUDA tensors only -- CPU tensors raise (no fallback).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import List, Optional, Sequence

import torch

from . import _lib

MODE_EVAL, MODE_GUMBEL, MODE_STE, MODE_ROTATION = 0, 1, 2, 3

#: count of librqb200 kernel launches issued through this module (bench.py reports it as `gpu_launches`)
LAUNCHES = 0
#: calls of the tensor-core tokeniser / prepared-state builds (tests assert the module API really routes there)
TC_CALLS = 0
TC_PREPARES = 0


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _p(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.Rqb200Error("rq_vae_recommender_b200 runs on CUDA tensors only (no CPU fallback); got a "
                                   f"{t.device} tensor")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _rows(t: torch.Tensor) -> torch.Tensor:
    """fp32 2-D tensor whose last dim is unit-stride (row stride may exceed the width)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _ptr_array(ts: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = _p(t)
    return arr


def _check_codebooks(cbs: Sequence[torch.Tensor], D: int) -> List[torch.Tensor]:
    out = [_f32c(c) for c in cbs]
    K = out[0].shape[0]
    for c in out:
        if c.shape != (K, D):
            raise ValueError(f"codebook shape {tuple(c.shape)} != ({K}, {D})")
    return out


# ---------------------------------------------------------------------------------------------- fused RQ chain
def rq_forward(x: torch.Tensor, codebooks: Sequence[torch.Tensor], mode: int, beta: float, *,
               want_ids=True, want_embeddings=False, want_residuals=False, want_sum=False, want_norms=False,
               want_loss=False):
    """All L Quantize levels in one launch (csrc/rq_simt.cu).  Returns a dict of the requested outputs;
    embeddings / residuals come back as [L,B,D] (permute to the reference's [B,D,L] is a view)."""
    _need_cuda(x, *codebooks)
    lib = _lib.load()
    x = _rows(x)
    B, D = x.shape
    cbs = _check_codebooks(codebooks, D)
    K, L = cbs[0].shape[0], len(cbs)
    dev = x.device
    out = {}
    ids = torch.empty((B, L), dtype=torch.int64, device=dev) if want_ids else None
    emb = torch.empty((L, B, D), dtype=torch.float32, device=dev) if want_embeddings else None
    res = torch.empty((L, B, D), dtype=torch.float32, device=dev) if want_residuals else None
    esum = torch.empty((B, D), dtype=torch.float32, device=dev) if want_sum else None
    norms = torch.empty((B, L), dtype=torch.float32, device=dev) if want_norms else None
    loss = torch.empty((B,), dtype=torch.float32, device=dev) if want_loss else None
    if (B >= TC_MIN_ROWS and tc_padded_dim(D, K, L)
            and (want_embeddings or want_residuals or want_sum or want_norms or want_loss)):
        # large batch: ids from the tensor-core tokeniser (those of the exact kernel), everything else from the streaming
        # pass over the given ids -- same outputs bit for bit, HBM-bound instead of CUDA-core-FLOP-bound
        tids = rq_tokenize_tc(x, state=tc_state_for(cbs))
        with torch.cuda.device(dev):
            _lib.check(lib.rqb200_rq_forward_from_ids(mode, _p(x), x.stride(0), _ptr_array(cbs), _p(tids), B, D, K, L,
                                                      float(beta), _p(emb), _p(res), _p(esum), _p(norms), _p(loss), _stream()),
                       "rq_forward_from_ids")
        _count(1)
        out.update(ids=tids if want_ids else None, embeddings=emb, residuals=res, emb_sum=esum, emb_norms=norms, loss=loss)
        return out
    ws_bytes = lib.rqb200_rq_workspace_bytes(D, K, L)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rqb200_rq_forward(mode, _p(x), x.stride(0), _ptr_array(cbs), B, D, K, L, float(beta),
                                         _p(ids), _p(emb), _p(res), _p(esum), _p(norms), _p(loss),
                                         _p(ws), ws_bytes, _stream()), "rq_forward")
    _count(3)
    out.update(ids=ids, embeddings=emb, residuals=res, emb_sum=esum, emb_norms=norms, loss=loss)
    return out


def rq_tokenize(x: torch.Tensor, codebooks: Sequence[torch.Tensor]) -> torch.Tensor:
    """sem_ids [B,L] int64 -- eval-mode hard argmin chain, exact fp32 kernel."""
    return rq_forward(x, codebooks, MODE_EVAL, 0.0, want_ids=True)["ids"]


class RqChainFunction(torch.autograd.Function):
    """Differentiable fused chain (eval / STE / rotation-trick).

    lean=True  -> (emb_sum [B,D], emb_norms [B,L], ids [B,L], loss [B])     what RqVae.forward consumes
    lean=False -> (embeddings [L,B,D], residuals [L,B,D], ids [B,L], loss [B])   what get_semantic_ids returns
    """

    @staticmethod
    def forward(ctx, x, mode, beta, lean, *codebooks):
        xr = _rows(x)
        cbs = _check_codebooks(codebooks, xr.shape[1])
        o = rq_forward(xr, cbs, mode, beta, want_ids=True, want_embeddings=not lean, want_residuals=not lean,
                       want_sum=lean, want_norms=lean, want_loss=True)
        ctx.mode, ctx.beta, ctx.lean = mode, beta, lean
        ctx.save_for_backward(xr, o["ids"], *cbs)
        ctx.mark_non_differentiable(o["ids"])
        if lean:
            ctx.mark_non_differentiable(o["emb_norms"])
            return o["emb_sum"], o["emb_norms"], o["ids"], o["loss"]
        return o["embeddings"], o["residuals"], o["ids"], o["loss"]

    @staticmethod
    def backward(ctx, g_a, g_b, _g_ids, g_loss):
        xr, ids, *cbs = ctx.saved_tensors
        lib = _lib.load()
        B, D = xr.shape
        K, L = cbs[0].shape[0], len(cbs)
        dev = xr.device

        def f32(g):
            return None if g is None else (g if g.dtype == torch.float32 else g.float())

        g_a, g_loss = f32(g_a), f32(g_loss)
        if ctx.lean:
            g_emb, g_res = g_a, None
            ge = (g_emb.stride(0), g_emb.stride(1), 0) if g_emb is not None else (0, 0, 0)
            gr = (0, 0, 0)
        else:
            g_emb, g_res = g_a, f32(g_b)
            ge = (g_emb.stride(1), g_emb.stride(2), g_emb.stride(0)) if g_emb is not None else (0, 0, 0)
            gr = (g_res.stride(1), g_res.stride(2), g_res.stride(0)) if g_res is not None else (0, 0, 0)
        g_x = torch.empty((B, D), dtype=torch.float32, device=dev)
        need_cb = [ctx.needs_input_grad[4 + l] for l in range(L)]
        g_cbs = [torch.zeros_like(c) if n else None for c, n in zip(cbs, need_cb)]
        with torch.cuda.device(dev):
            _lib.check(lib.rqb200_rq_backward(ctx.mode, _p(xr), xr.stride(0), _ptr_array(cbs), _p(ids), B, D, K, L,
                                              float(ctx.beta), _p(g_emb), *ge, _p(g_res), *gr, _p(g_loss),
                                              g_loss.stride(0) if g_loss is not None else 0, _p(g_x),
                                              _ptr_array(g_cbs), _stream()), "rq_backward")
        _count(1)
        return (g_x if ctx.needs_input_grad[0] else None, None, None, None, *g_cbs)


# ---------------------------------------------------------------------------------------------- dense helpers
def sgemm(a: torch.Tensor, b: torch.Tensor, *, trans_a=False, trans_b=False, relu=False,
          mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, alpha=1.0, beta=0.0):
    """out = epi(alpha * op(a) @ op(b) + beta*out) with the fp32 CUDA-core GEMM of csrc/dense.cu."""
    lib = _lib.load()
    a, b = _rows(a), _rows(b)
    M, Ka = (a.shape[1], a.shape[0]) if trans_a else a.shape
    Kb, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if Ka != Kb:
        raise ValueError(f"sgemm inner dims differ: {Ka} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if mask is not None:
        mask = _rows(mask)
    with torch.cuda.device(a.device):
        _lib.check(lib.rqb200_sgemm(int(trans_a), int(trans_b), M, N, Ka, float(alpha), _p(a), a.stride(0), _p(b),
                                    b.stride(0), float(beta), _p(out), out.stride(0), int(relu), _p(mask),
                                    mask.stride(0) if mask is not None else 0, _stream()), "sgemm")
    _count(1)
    return out


# ---------------------------------------------------------------------------------------------- split-precision tensor-core GEMM
#: rows of the A operand from which the dense GEMMs of the MLPs / the Gumbel level run on the tensor cores (three fp16
#: products per k-step, fp32-accurate: csrc/gemm_tc.cu); below it one launch of the CUDA-core SGEMM is as fast
SPLIT_MIN_ROWS = 512
SPLIT_CALLS = 0


class SplitOperand:
    """One GEMM operand [rows, K] as hi + lo fp16 images with power-of-two row scales (rqb200_f32_to_split_image).
    ``transposed=True`` builds the operand of t.T without materialising the transpose."""

    def __init__(self, t: torch.Tensor, transposed: bool = False):
        _need_cuda(t)
        lib = _lib.load()
        t = _rows(t.detach())
        self.rows, self.K = (t.shape[1], t.shape[0]) if transposed else (t.shape[0], t.shape[1])
        self.device = t.device
        nbytes = lib.rqb200_split_image_bytes(self.rows, self.K)
        self.buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(lib.rqb200_f32_to_split_image(_p(t), t.stride(0), self.rows, self.K, int(transposed), _p(self.buf),
                                                     _stream()), "f32_to_split_image")
        _count(1)


# operands that do not change between calls (weights, codebooks): keyed on identity AND version like the tokeniser state
_SPLIT_CACHE: "dict[tuple, tuple]" = {}
_SPLIT_CACHE_MAX = 32


_NO_OPERAND_CACHE = 0


class no_operand_cache:
    """Inside this context prepared operands are rebuilt on every use and never stored.  The torch.library operators run under
    it: a compiled caller (mode="reduce-overhead") executes them inside a CUDA-graph memory pool -- warm-up run included -- where
    a cached buffer would be an allocation the graph does not own, and a replay must see the weights of that moment."""

    def __enter__(self):
        global _NO_OPERAND_CACHE
        _NO_OPERAND_CACHE += 1

    def __exit__(self, *a):
        global _NO_OPERAND_CACHE
        _NO_OPERAND_CACHE -= 1


def split_operand_cached(t: torch.Tensor, transposed: bool = False) -> SplitOperand:
    # the entry remembers the tensor OBJECT (weak reference): a temporary that died and whose address the allocator handed to
    # another tensor of the same shape must not hit
    if _NO_OPERAND_CACHE or torch.cuda.is_current_stream_capturing():
        return SplitOperand(t, transposed)       # inside a CUDA-graph capture the split kernel must be PART of the graph (replays
                                                 # see the weights of that moment, not the images of capture time)
    key = (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.device.index, bool(transposed))
    hit = _SPLIT_CACHE.get(key)
    if hit is not None and hit[0]() is t:
        return hit[1]
    if len(_SPLIT_CACHE) >= _SPLIT_CACHE_MAX:
        _SPLIT_CACHE.pop(next(iter(_SPLIT_CACHE)))
    op = SplitOperand(t, transposed)
    _SPLIT_CACHE[key] = (weakref.ref(t), op)
    return op


def gemm_split(a, b, *, relu: bool = False, mask: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M, N] = act(A[M, K] @ B[N, K]^T) on the fp16 tensor cores with fp32 accuracy; a, b: SplitOperand or fp32 tensor."""
    global SPLIT_CALLS
    lib = _lib.load()
    if not isinstance(a, SplitOperand):
        a = SplitOperand(a)
    if not isinstance(b, SplitOperand):
        b = SplitOperand(b)
    if a.K != b.K:
        raise ValueError(f"gemm_split inner dims differ: {a.K} vs {b.K}")
    M, N = a.rows, b.rows
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if mask is not None:
        mask = _rows(mask)
    with torch.cuda.device(a.device):
        _lib.check(lib.rqb200_gemm_split(_p(a.buf), _p(b.buf), M, N, a.K, int(relu), _p(mask),
                                         mask.stride(0) if mask is not None else 0, _p(out), out.stride(0), _stream()),
                   "gemm_split")
    _count(1)
    SPLIT_CALLS += 1
    return out


def gemm_tn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a^T @ b for a [B, M], b [B, N]: the contraction runs over the (long) batch dimension -- weight / codebook gradients.
    Tensor cores with a split-K schedule from SPLIT_MIN_ROWS batch rows on (both operands are split transposed, partial sums are
    reduced in a fixed order), the CUDA-core SGEMM below."""
    global SPLIT_CALLS
    B = a.shape[0]
    if B < SPLIT_MIN_ROWS:
        return sgemm(a, b, trans_a=True)
    lib = _lib.load()
    ao, bo = SplitOperand(a, transposed=True), SplitOperand(b, transposed=True)
    M, N = ao.rows, bo.rows
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        slices = lib.rqb200_gemm_split_k_slices(M, N, B)
        ws = torch.empty((slices, M, N), dtype=torch.float32, device=a.device) if slices > 1 else None
        _lib.check(lib.rqb200_gemm_split_k(_p(ao.buf), _p(bo.buf), M, N, B, slices, _p(ws), _p(out), out.stride(0), _stream()),
                   "gemm_split_k")
    _count(2 if slices > 1 else 1)
    SPLIT_CALLS += 1
    return out


def linear_nt(a: torch.Tensor, w: torch.Tensor, *, relu: bool = False, mask: Optional[torch.Tensor] = None,
              w_transposed: bool = False) -> torch.Tensor:
    """a[M, K] @ op(w)^T with a static second operand (weight / codebook): op(w) = w [N, K], or w^T for w [K, N] when
    ``w_transposed``.  Tensor cores from SPLIT_MIN_ROWS rows on, the CUDA-core SGEMM below (same result to ~1e-7)."""
    if a.shape[0] >= SPLIT_MIN_ROWS:
        return gemm_split(a, split_operand_cached(w, transposed=w_transposed), relu=relu, mask=mask)
    return sgemm(a, w, trans_b=not w_transposed, relu=relu, mask=mask)


class MLPFunction(torch.autograd.Function):
    """modules/encoder.py:23-38 as one autograd node: bias-free Linear+ReLU stack (ReLU fused in the GEMM
    epilogue), optional final L2 normalisation (modules/normalize.py)."""

    @staticmethod
    def forward(ctx, x, normalize, *weights):
        _need_cuda(x, *weights)
        lib = _lib.load()
        h = _rows(x)
        ws = [_f32c(w) for w in weights]
        acts = [h]
        n = len(ws)
        for i, w in enumerate(ws):
            h = linear_nt(h, w, relu=(i != n - 1))
            acts.append(h)
        norms = None
        if normalize:
            y = torch.empty_like(h)
            norms = torch.empty(h.shape[0], dtype=torch.float32, device=h.device)
            with torch.cuda.device(h.device):
                _lib.check(lib.rqb200_l2norm_fwd(_p(h), _p(y), _p(norms), h.shape[0], h.shape[1], 1e-12, _stream()),
                           "l2norm_fwd")
            _count(1)
            acts.append(y)
            h = y
        ctx.normalize = normalize
        ctx.n = n
        ctx.save_for_backward(*acts, *ws, *([norms] if normalize else []))
        return h

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        n = ctx.n
        saved = ctx.saved_tensors
        n_act = n + 1 + (1 if ctx.normalize else 0)
        acts, ws = saved[:n_act], saved[n_act:n_act + n]
        g = _f32c(g)
        if ctx.normalize:
            norms, y = saved[-1], acts[-1]
            gx = torch.empty_like(g)
            with torch.cuda.device(g.device):
                _lib.check(lib.rqb200_l2norm_bwd(_p(g), _p(y), _p(norms), _p(gx), g.shape[0], g.shape[1], 1e-12,
                                                 _stream()), "l2norm_bwd")
            _count(1)
            g = gx
        g_ws = [None] * n
        for i in range(n - 1, -1, -1):
            h_in = acts[i]
            if ctx.needs_input_grad[2 + i]:
                g_ws[i] = gemm_tn(g, h_in)                                 # [out,B] @ [B,in]
            if i > 0 or ctx.needs_input_grad[0]:
                g = linear_nt(g, ws[i], mask=h_in if i > 0 else None, w_transposed=True)   # [B,out] @ [out,in], ReLU' of layer i-1
        return (g if ctx.needs_input_grad[0] else None, None, *g_ws)


def l2norm_rows(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, p=2, dim=-1) forward only (no autograd)."""
    lib = _lib.load()
    shp = x.shape
    x2 = _f32c(x.reshape(-1, shp[-1]))
    y = torch.empty_like(x2)
    with torch.cuda.device(x2.device):
        _lib.check(lib.rqb200_l2norm_fwd(_p(x2), _p(y), 0, x2.shape[0], x2.shape[1], float(eps), _stream()), "l2norm")
    _count(1)
    return y.reshape(shp)


class L2NormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        lib = _lib.load()
        shp = x.shape
        x2 = _f32c(x.reshape(-1, shp[-1]))
        y = torch.empty_like(x2)
        norms = torch.empty(x2.shape[0], dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            _lib.check(lib.rqb200_l2norm_fwd(_p(x2), _p(y), _p(norms), x2.shape[0], x2.shape[1], float(eps), _stream()),
                       "l2norm_fwd")
        _count(1)
        ctx.eps, ctx.shp = eps, shp
        ctx.save_for_backward(y, norms)
        return y.reshape(shp)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        y, norms = ctx.saved_tensors
        g2 = _f32c(g.reshape(y.shape))
        gx = torch.empty_like(g2)
        with torch.cuda.device(g2.device):
            _lib.check(lib.rqb200_l2norm_bwd(_p(g2), _p(y), _p(norms), _p(gx), y.shape[0], y.shape[1], float(ctx.eps),
                                             _stream()), "l2norm_bwd")
        _count(1)
        return gx.reshape(ctx.shp), None


# ---------------------------------------------------------------------------------------------- Gumbel-softmax level
class GumbelQuantizeFunction(torch.autograd.Function):
    """One training-mode GUMBEL_SOFTMAX Quantize level (modules/quantize.py:113-136,157): returns
    (emb [B,D], ids [B], loss [B]).  The uniform draw U is an input (the caller draws it with torch.rand on the
    device, exactly where distributions/gumbel.py:10 does)."""

    @staticmethod
    def forward(ctx, x, codebook, uniform, temperature, beta):
        _need_cuda(x, codebook, uniform)
        lib = _lib.load()
        x, cb, u = _rows(x), _f32c(codebook), _f32c(uniform)
        B, D = x.shape
        K = cb.shape[0]
        dev = x.device
        st = _stream()
        cc = torch.empty(K, dtype=torch.float32, device=dev)
        ids = torch.empty(B, dtype=torch.int64, device=dev)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        w = torch.empty((B, K), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.rqb200_row_sqnorm(_p(cb), K, D, _p(cc), st), "row_sqnorm")
            dist = linear_nt(x, cb)                                              # x @ C^T
            _lib.check(lib.rqb200_dist_finish(_p(dist), _p(x), x.stride(0), _p(cc), B, D, K, _p(ids), st), "dist")
            _lib.check(lib.rqb200_gumbel_softmax_fwd(_p(dist), _p(u), _p(w), B, K, float(temperature), st), "softmax")
            emb = linear_nt(w, cb, w_transposed=True)                            # W @ C   (quantize.py:135)
            _lib.check(lib.rqb200_gumbel_row_finish(_p(x), x.stride(0), _p(emb), B, D, float(beta), _p(loss), st),
                       "row_finish")
        _count(4)
        ctx.temperature, ctx.beta = float(temperature), float(beta)
        ctx.save_for_backward(x, cb, w, emb)
        ctx.mark_non_differentiable(ids)
        return emb, ids, loss

    @staticmethod
    def backward(ctx, g_emb, _g_ids, g_loss):
        lib = _lib.load()
        x, cb, w, emb = ctx.saved_tensors
        B, D = x.shape
        K = cb.shape[0]
        dev = x.device
        st = _stream()
        g_emb = None if g_emb is None else (g_emb if g_emb.dtype == torch.float32 else g_emb.float())
        g_loss = None if g_loss is None else (g_loss if g_loss.dtype == torch.float32 else g_loss.float())
        gl_s = g_loss.stride(0) if g_loss is not None else 0
        gE = torch.empty((B, D), dtype=torch.float32, device=dev)
        rowsum = torch.empty(B, dtype=torch.float32, device=dev)
        colsum = torch.empty(K, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.rqb200_gumbel_bwd_ge(_p(g_emb), g_emb.stride(0) if g_emb is not None else 0,
                                                g_emb.stride(1) if g_emb is not None else 0, _p(g_loss), gl_s,
                                                _p(x), x.stride(0), _p(emb), _p(gE), B, D, st), "bwd_ge")
            gw = linear_nt(gE, cb)                                               # gW = gE @ C^T   [B,K]
            _lib.check(lib.rqb200_gumbel_bwd_softmax(_p(w), _p(gw), B, K, ctx.temperature, _p(rowsum), _p(colsum), st),
                       "bwd_softmax")                                            # gw now holds gdist
            gx = linear_nt(gw, cb, w_transposed=True)                            # gdist @ C       [B,D]
            _lib.check(lib.rqb200_gumbel_bwd_gx(_p(gx), _p(x), x.stride(0), _p(emb), _p(g_loss), gl_s, _p(rowsum),
                                                ctx.beta, B, D, st), "bwd_gx")
            gc = None
            if ctx.needs_input_grad[1]:
                if B >= SPLIT_MIN_ROWS:
                    gc = gemm_tn(w, gE)                                          # W^T @ gE        [K,D]
                    gc.add_(gemm_tn(gw, x), alpha=-2.0)                          # - 2 gdist^T @ x
                else:
                    gc = sgemm(w, gE, trans_a=True)
                    sgemm(gw, x, trans_a=True, out=gc, alpha=-2.0, beta=1.0)
                _lib.check(lib.rqb200_gumbel_bwd_gc(_p(gc), _p(cb), _p(colsum), K, D, st), "bwd_gc")
        _count(5)
        return (gx if ctx.needs_input_grad[0] else None, gc, None, None, None)


# ---------------------------------------------------------------------------------------------- k-means
def kmeans_workspace(x: torch.Tensor, K: int):
    lib = _lib.load()
    D = x.shape[1]
    dev = x.device
    ws_bytes = lib.rqb200_rq_workspace_bytes(D, K, 1)
    return dict(ws=torch.empty(ws_bytes, dtype=torch.uint8, device=dev), ws_bytes=ws_bytes,
                assign=torch.empty(x.shape[0], dtype=torch.int64, device=dev),
                sums=torch.empty((K, D), dtype=torch.float64, device=dev),
                counts=torch.empty(K, dtype=torch.int32, device=dev),
                shift=torch.empty(1, dtype=torch.float32, device=dev))


def kmeans_assign_accumulate(x: torch.Tensor, centroids: torch.Tensor, buf) -> None:
    lib = _lib.load()
    B, D = x.shape
    K = centroids.shape[0]
    with torch.cuda.device(x.device):
        _lib.check(lib.rqb200_kmeans_assign_accumulate(_p(x), x.stride(0), _p(centroids), B, D, K, _p(buf["assign"]),
                                                       _p(buf["sums"]), _p(buf["counts"]), _p(buf["ws"]),
                                                       buf["ws_bytes"], _stream()), "kmeans_assign_accumulate")
    _count(3)


def kmeans_finalize(x: torch.Tensor, centroids: torch.Tensor, buf, reseed_rows: Optional[torch.Tensor]) -> None:
    lib = _lib.load()
    K, D = centroids.shape
    with torch.cuda.device(x.device):
        _lib.check(lib.rqb200_kmeans_finalize(_p(buf["sums"]), _p(buf["counts"]), _p(x), x.stride(0), _p(reseed_rows),
                                              _p(centroids), K, D, _p(buf["shift"]), _stream()), "kmeans_finalize")
    _count(1)


# ---------------------------------------------------------------------------------------------- id statistics
def sid_histogram(ids: torch.Tensor, K: int) -> torch.Tensor:
    """[L,K] int64 code-usage counts of a [B,L] int64 id table (train_rqvae.py:285-289)."""
    _need_cuda(ids)
    lib = _lib.load()
    ids = ids.contiguous()
    B, L = ids.shape
    hist = torch.empty((L, K), dtype=torch.int64, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.check(lib.rqb200_sid_histogram(_p(ids), B, L, K, _p(hist), _stream()), "sid_histogram")
    _count(1)
    return hist


def sid_dedup_rank(ids: torch.Tensor, K: int):
    """Dedup column + diversity statistics of a corpus id table (semids.py:94-108, train_rqvae.py:276-283) in two launches.
    Returns (rank [N] int64, stats) with stats = dict(max_rank, n_unique: 0-d int32 tensors, entropy: 0-d float64 tensor), all on
    the device (no host sync), or None when the key space K^L is too large for the direct table (the caller sorts instead)."""
    _need_cuda(ids)
    lib = _lib.load()
    ids = ids.contiguous()
    N, L = ids.shape
    ws_bytes = lib.rqb200_sid_dedup_workspace_bytes(N, L, K)
    if ws_bytes == 0:
        return None
    dev = ids.device
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rank = torch.empty(N, dtype=torch.int64, device=dev)
    stats = torch.empty(2, dtype=torch.int32, device=dev)
    entropy = torch.empty(1, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rqb200_sid_dedup_rank(_p(ids), N, L, K, _p(rank), _p(stats), _p(entropy), _p(ws), ws_bytes, _stream()),
                   "sid_dedup_rank")
    _count(2)
    return rank, dict(max_rank=stats[0], n_unique=stats[1], entropy=entropy[0])


class SidPrefixIndex:
    """Valid-prefix index of a corpus id table [N, C] (modules/model.py:169-182): one bitmap per prefix length, built once per
    corpus by rqb200_sid_prefix_build.  ``check`` is the reference's `_check_valid_prefix`; ``beam_select`` one selection step of
    its constrained beam search (model.py:340-376)."""

    def __init__(self, cached_ids: torch.Tensor, codebook_size: int):
        _need_cuda(cached_ids)
        lib = _lib.load()
        ids = cached_ids.to(torch.int64).contiguous()
        self.N, self.C = ids.shape
        self.K = int(codebook_size)
        self.device = ids.device
        nbytes = lib.rqb200_sid_prefix_workspace_bytes(self.C, self.K)
        if nbytes == 0:
            raise _lib.Rqb200Error(f"prefix index: key space {self.K}^{self.C} exceeds the bitmap limit (2^33 bits)")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=ids.device)
        with torch.cuda.device(ids.device):
            _lib.check(lib.rqb200_sid_prefix_build(_p(ids), self.N, self.C, self.K, _p(self.ws), nbytes, _stream()),
                       "sid_prefix_build")
        _count(1)

    def check(self, prefix: torch.Tensor) -> torch.Tensor:
        """bool [P]: does some corpus row start with prefix[p] ([P, l], l <= C)."""
        _need_cuda(prefix)
        lib = _lib.load()
        if prefix.dtype != torch.int64:
            prefix = prefix.to(torch.int64)
        if prefix.stride(-1) != 1:
            prefix = prefix.contiguous()
        P, l = prefix.shape
        if l > self.C:
            raise ValueError(f"prefix length {l} exceeds the id tuple length {self.C}")
        valid = torch.empty(P, dtype=torch.bool, device=prefix.device)
        with torch.cuda.device(prefix.device):
            _lib.check(lib.rqb200_sid_prefix_check(_p(prefix), prefix.stride(0), P, l, self.C, self.K, _p(self.ws), _p(valid),
                                                   _stream()), "sid_prefix_check")
        _count(1)
        return valid

    def beam_select(self, samples: torch.Tensor, samp_log_p: torch.Tensor, generated: Optional[torch.Tensor],
                    log_probas: Optional[torch.Tensor], k: int):
        """samples / samp_log_p [B * kp, nc] (kp = 1 on the first level), generated [B, kp, h] or None, log_probas [B, kp] or None
        -> (generated [B, k, h + 1], log_probas [B, k], parent_global [B * k]) exactly as model.py:353-388 computes them."""
        _need_cuda(samples, samp_log_p)
        lib = _lib.load()
        if generated is None:
            B, kp, h = samples.shape[0], 1, 0
        else:
            B, kp, h = generated.shape
            generated = generated.to(torch.int64).contiguous()
        nc = samples.shape[-1]
        samples = samples.to(torch.int64).reshape(B * kp, nc).contiguous()
        samp_log_p = samp_log_p.to(torch.float32).reshape(B * kp, nc).contiguous()
        if log_probas is not None:
            log_probas = log_probas.to(torch.float32).reshape(B, kp).contiguous()
        dev = samples.device
        out_g = torch.empty((B, k, h + 1), dtype=torch.int64, device=dev)
        out_p = torch.empty((B, k), dtype=torch.float32, device=dev)
        out_parent = torch.empty((B * k,), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.rqb200_sid_beam_select(_p(samples), _p(samp_log_p), _p(generated), _p(log_probas), B, kp, nc, h, k,
                                                  self.C, self.K, _p(self.ws), _p(out_g), _p(out_p), _p(out_parent), _stream()),
                       "sid_beam_select")
        _count(1)
        return out_g, out_p, out_parent


def sid_gather(cached_ids: torch.Tensor, item_ids: torch.Tensor, seq_mask: Optional[torch.Tensor] = None,
               want_token_type: bool = True):
    """cached_ids[item_ids] -> [B, S*C] with -1 under the padding mask, and token_type_ids (semids.py:112-146), one launch."""
    _need_cuda(cached_ids, item_ids)
    lib = _lib.load()
    cached_ids = cached_ids.contiguous()
    if item_ids.stride(-1) != 1:
        item_ids = item_ids.contiguous()
    B, S = item_ids.shape
    C = cached_ids.shape[1]
    dev = cached_ids.device
    out = torch.empty((B, S * C), dtype=torch.int64, device=dev)
    tt = torch.empty((B, S * C), dtype=torch.int64, device=dev) if want_token_type else None
    m = None
    if seq_mask is not None:
        m = seq_mask.to(torch.uint8) if seq_mask.dtype != torch.uint8 else seq_mask
        if m.stride(-1) != 1:
            m = m.contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.rqb200_sid_gather(_p(cached_ids), cached_ids.shape[0], C, _p(item_ids), item_ids.stride(0), _p(m),
                                         m.stride(0) if m is not None else 0, B, S, _p(out), _p(tt), _stream()), "sid_gather")
    _count(1)
    return out, tt


# ---------------------------------------------------------------------------------------------- tensor-core tokeniser
def tc_supported(D: int, K: int, L: int) -> bool:
    return bool(_lib.load().rqb200_tokenize_tc_supported(D, K, L))


TC_PAD = 64   # the tcgen05 tokeniser wants D % 64 == 0: narrower / odd widths are zero-padded (exact for every dot product)


def tc_padded_dim(D: int, K: int, L: int) -> int:
    """Width the tensor-core tokeniser runs a D-wide quantiser at (D itself, or D zero-padded to the next multiple of 64);
    0 when the shape cannot use it at all (K != 256, D > 768, L > 8)."""
    Dp = (D + TC_PAD - 1) // TC_PAD * TC_PAD
    return Dp if tc_supported(Dp, K, L) else 0


class TcState:
    """Device-side prepared codebooks for the tcgen05 tokeniser (fp16 images, measured rounding norms, float64 Gram
    tables, an fp32 copy for the exact re-rank).  Owns everything it needs: the caller's tensors are not referenced after
    construction.  Codebooks narrower than a multiple of 64 are zero-padded (``self.D`` is the padded width,
    ``self.D_in`` the caller's)."""

    def __init__(self, codebooks: Sequence[torch.Tensor]):
        lib = _lib.load()
        cbs = _check_codebooks(codebooks, codebooks[0].shape[1])
        _need_cuda(*cbs)
        self.K, self.D_in = cbs[0].shape
        self.L = len(cbs)
        self.D = tc_padded_dim(self.D_in, self.K, self.L)
        if not self.D:
            raise _lib.Rqb200Error(f"tcgen05 tokeniser does not support D={self.D_in} K={self.K} L={self.L}")
        if self.D != self.D_in:
            cbs = [torch.nn.functional.pad(c, (0, self.D - self.D_in)) for c in cbs]
        self.device = cbs[0].device
        nbytes = lib.rqb200_tokenize_tc_state_bytes(self.D, self.K, self.L)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.nbytes = nbytes
        with torch.cuda.device(self.device):
            _lib.check(lib.rqb200_tokenize_tc_prepare(_ptr_array(cbs), self.D, self.K, self.L, _p(self.buf), nbytes,
                                                      _stream()), "tokenize_tc_prepare")
        _count(6 + self.L * (self.L - 1) // 2)
        global TC_PREPARES
        TC_PREPARES += 1


def rq_tokenize_tc(x: torch.Tensor, codebooks=None, state: Optional[TcState] = None, stats=None) -> torch.Tensor:
    """sem_ids [B,L] int64 via the tcgen05 candidate filter + exact fp32 re-rank (csrc/rq_tcx.cu).  Same result contract as
    ``rq_tokenize``: the filter's margin is a deterministic bound on the fp16 rounding, every row with more
    than one candidate inside it is re-scored with the exact kernel's fp32 arithmetic."""
    _need_cuda(x)
    lib = _lib.load()
    if state is None:
        state = TcState(codebooks)
    x = _rows(x)
    B, D = x.shape
    if D != state.D_in:
        raise _lib.Rqb200Error(f"rq_tokenize_tc: x has {D} columns, the prepared state was built for {state.D_in}")
    if x.device != state.device:
        raise _lib.Rqb200Error(f"rq_tokenize_tc: x is on {x.device}, the prepared state on {state.device}")
    if D != state.D:
        x = torch.nn.functional.pad(x, (0, state.D - D))
    if x.data_ptr() % 16 or x.stride(0) % 4:        # the kernel reads x through TMA: 16-byte aligned base and row pitch
        x = x.contiguous()
    ids = torch.empty((B, state.L), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.rqb200_tokenize_tc_run(_p(x), x.stride(0), B, _p(state.buf), state.D, state.K, state.L, _p(ids),
                                              _p(stats), _stream()), "tokenize_tc_run")
    _count(1)
    global TC_CALLS
    TC_CALLS += 1
    return ids


# frozen-codebook cache of prepared states behind the module API (RqVae.tokenize / SemanticIdTokenizer): keyed on the
# identity AND version of every codebook tensor, so an optimiser step (in-place update bumps ._version) or a reloaded
# checkpoint re-prepares; a handful of entries is plenty (one model per process in the reference's scripts)
_TC_CACHE: "dict[tuple, tuple]" = {}
_TC_CACHE_MAX = 4
#: rows below which the exact CUDA-core kernel is used even when the tensor-core path is available (one 128-row tile keeps
#: 2 of 148 SMs busy; the prepare step costs ~0.4 ms when the cache misses)
TC_MIN_ROWS = 1024


def _tc_cache_key(codebooks: Sequence[torch.Tensor]):
    return tuple((c.data_ptr(), c._version, tuple(c.shape), c.device.index) for c in codebooks)


def tc_state_for(codebooks: Sequence[torch.Tensor]) -> TcState:
    # an entry also remembers the tensor OBJECTS (weak references): a derived codebook (sim_vq projection, normalised rows) is a
    # temporary whose address the allocator may hand to the next call's temporary with the same version counter
    if _NO_OPERAND_CACHE or torch.cuda.is_current_stream_capturing():
        return TcState(codebooks)                # a captured graph re-prepares on every replay (see split_operand_cached)
    key = _tc_cache_key(codebooks)
    hit = _TC_CACHE.get(key)
    if hit is not None and all(r() is c for r, c in zip(hit[0], codebooks)):
        return hit[1]
    if len(_TC_CACHE) >= _TC_CACHE_MAX:
        _TC_CACHE.pop(next(iter(_TC_CACHE)))
    st = TcState(codebooks)
    _TC_CACHE[key] = ([weakref.ref(c) for c in codebooks], st)
    return st


def rq_tokenize_auto(x: torch.Tensor, codebooks: Sequence[torch.Tensor], stats=None) -> torch.Tensor:
    """What the module API calls (RqVae.tokenize, SemanticIdTokenizer.precompute_corpus_ids, modules/rqvae.py:118-139 ids
    only): the tensor-core tokeniser with a cached prepared state whenever the shape allows (K = 256, D <= 768 after
    zero-padding to a multiple of 64) and the batch is large enough to fill the GPU, else the exact CUDA-core kernel."""
    K, D = codebooks[0].shape
    if x.shape[0] >= TC_MIN_ROWS and not torch.is_grad_enabled() and tc_padded_dim(D, K, len(codebooks)):
        return rq_tokenize_tc(x, state=tc_state_for(codebooks), stats=stats)
    return rq_tokenize(x, codebooks)


# ---------------------------------------------------------------------------------------------- bf16 tensor-core MLP
def bf16_supported(dims) -> bool:
    """Every contraction dim (all but the last entry of [in, hidden..., out]) must be a multiple of 64."""
    return all(d % 64 == 0 for d in dims[:-1])


def to_bf16_image(x: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, K] -> bf16 operand image (csrc/gemm_tc.cu); used for activations and for weights W[N, K]."""
    _need_cuda(x)
    lib = _lib.load()
    x = _rows(x)
    rows, K = x.shape
    nbytes = lib.rqb200_bf16_image_bytes(rows, K)
    if nbytes == 0 and rows:
        raise _lib.Rqb200Error(f"bf16 image needs K % 64 == 0, got K={K}")
    img = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.rqb200_f32_to_bf16_image(_p(x), x.stride(0), rows, K, _p(img), _stream()), "f32_to_bf16_image")
    _count(1)
    return img


@torch.no_grad()
def mlp_forward_bf16(x: torch.Tensor, weights: Sequence[torch.Tensor], normalize: bool = False,
                     weight_images: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
    """modules/encoder.py:23-38 with bf16 tcgen05 GEMMs (fp32 accumulate, ReLU fused; activations stay in the bf16
    operand-image layout between layers).  Forward only, reduced precision: NOT index-exact vs the fp32 reference."""
    _need_cuda(x, *weights)
    lib = _lib.load()
    x = _rows(x)
    M = x.shape[0]
    dims = [x.shape[1]] + [w.shape[0] for w in weights]
    if not bf16_supported(dims):
        raise _lib.Rqb200Error(f"bf16 MLP needs every contraction dim to be a multiple of 64, got {dims}")
    if weight_images is None:
        weight_images = [to_bf16_image(w.detach()) for w in weights]
    a = to_bf16_image(x)
    n = len(weights)
    out = None
    with torch.cuda.device(x.device):
        for i, (w, wi) in enumerate(zip(weights, weight_images)):
            N, K = w.shape
            last = i == n - 1
            nxt = None
            if not last:
                nxt = torch.empty(lib.rqb200_bf16_image_bytes(M, N), dtype=torch.uint8, device=x.device)
            else:
                out = torch.empty((M, N), dtype=torch.float32, device=x.device)
            _lib.check(lib.rqb200_gemm_bf16(_p(a), _p(wi), M, N, K, int(not last), _p(nxt), _p(out) if last else 0,
                                            N if last else 0, _stream()), "gemm_bf16")
            _count(1)
            a = nxt
    return l2norm_rows(out) if normalize else out
