"""torch.library registration of the librqb200 entry points that sit on the reference's compiled path.

The reference decorates ``RqVae.forward`` with ``@torch.compile(mode="reduce-overhead")`` (modules/rqvae.py:141).  The kernels
here are reached through ctypes, which Dynamo cannot trace; registered as custom operators (namespace ``rqb200``) with fake
(shape-only) implementations and autograd formulas they become single nodes of the captured graph instead of graph breaks.
Each operator wraps the forward / backward of the autograd Function that the eager path uses (ops.py), so both paths run the
same code and the same kernels.  Eager callers keep using the autograd Functions directly (no dispatcher overhead); the modules
switch to these operators only while ``torch.compiler.is_compiling()``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import Tensor
from torch.library import custom_op

from . import ops


class _Ctx:
    """Stand-in for the autograd context object when an autograd Function's static methods are called directly."""

    def __init__(self, needs_input_grad=()):
        self.saved_tensors = ()
        self.needs_input_grad = tuple(needs_input_grad)

    def save_for_backward(self, *ts):
        self.saved_tensors = ts

    def mark_non_differentiable(self, *ts):
        pass


def _fresh(t: Tensor, *inputs: Tensor) -> Tensor:
    """Operator outputs may not alias operator inputs (a marshalling helper returns its argument when it is already fp32 and
    contiguous)."""
    for i in inputs:
        if t is i or t.data_ptr() == i.data_ptr():
            return t.clone()
    return t


# ------------------------------------------------------------------------------------------------------------------ MLP
@custom_op("rqb200::mlp_fwd", mutates_args=(), device_types="cuda")
def mlp_fwd(x: Tensor, normalize: bool, weights: Sequence[Tensor]) -> List[Tensor]:
    """[activations of every layer (the last one is the output), row norms when normalize] -- modules/encoder.py:23-38"""
    ctx = _Ctx()
    with ops.no_operand_cache():
        ops.MLPFunction.forward(ctx, x, normalize, *weights)
    n = len(weights)
    n_act = n + 1 + (1 if normalize else 0)
    acts = ctx.saved_tensors[1:n_act]                       # acts[0] is the input itself
    extra = [ctx.saved_tensors[-1]] if normalize else []
    return [_fresh(t, x, *weights) for t in (*acts, *extra)]


@mlp_fwd.register_fake
def _(x, normalize, weights):
    B = x.shape[0]
    outs = [x.new_empty((B, w.shape[0]), dtype=torch.float32) for w in weights]
    if normalize:
        outs.append(x.new_empty((B, weights[-1].shape[0]), dtype=torch.float32))
        outs.append(x.new_empty((B,), dtype=torch.float32))
    return outs


@custom_op("rqb200::mlp_bwd", mutates_args=(), device_types="cuda")
def mlp_bwd(g: Tensor, x: Tensor, normalize: bool, weights: Sequence[Tensor], saved: Sequence[Tensor],
            need_x: bool, need_w: Sequence[bool]) -> List[Tensor]:
    n = len(weights)
    ctx = _Ctx((need_x, False, *need_w))
    ctx.normalize, ctx.n = normalize, n
    x2 = ops._rows(x)
    ws = [ops._f32c(w) for w in weights]
    acts = [x2, *saved[:n + (1 if normalize else 0)]]
    ctx.saved_tensors = (*acts, *ws, *([saved[-1]] if normalize else []))
    with ops.no_operand_cache():
        gx, _none, *gws = ops.MLPFunction.backward(ctx, g)
    out = [gx if gx is not None else x.new_zeros(x.shape, dtype=torch.float32)]
    out += [gw if gw is not None else w.new_zeros(w.shape, dtype=torch.float32) for gw, w in zip(gws, weights)]
    return [_fresh(t, g, x, *weights, *saved) for t in out]


@mlp_bwd.register_fake
def _(g, x, normalize, weights, saved, need_x, need_w):
    return [x.new_empty(x.shape, dtype=torch.float32)] + [w.new_empty(w.shape, dtype=torch.float32) for w in weights]


def _mlp_setup(ctx, inputs, output):
    x, normalize, weights = inputs
    ctx.normalize = normalize
    ctx.n = len(weights)
    ctx.save_for_backward(x, *weights, *output)


def _mlp_backward(ctx, grads):
    n = ctx.n
    x, *rest = ctx.saved_tensors
    weights, saved = rest[:n], rest[n:]
    # the gradient of the OUTPUT: the last activation (or the normalised rows); the other returned tensors are saved state
    g = grads[n if ctx.normalize else n - 1]
    if g is None:
        return None, None, [None] * n
    need = ctx.needs_input_grad
    need_w = [True] * n                                     # per-element flags of a list input are not exposed: compute all
    out = mlp_bwd(g.contiguous(), x, ctx.normalize, list(weights), list(saved), bool(need[0]), need_w)
    return (out[0] if need[0] else None), None, list(out[1:])


mlp_fwd.register_autograd(_mlp_backward, setup_context=_mlp_setup)


def mlp(x: Tensor, normalize: bool, weights: Sequence[Tensor]) -> Tensor:
    outs = mlp_fwd(x, normalize, list(weights))
    return outs[len(weights) if normalize else len(weights) - 1]


# ------------------------------------------------------------------------------------------------------------ fused RQ chain
@custom_op("rqb200::rq_chain_fwd", mutates_args=(), device_types="cuda")
def rq_chain_fwd(x: Tensor, mode: int, beta: float, lean: bool, codebooks: Sequence[Tensor]) -> List[Tensor]:
    """lean: [emb_sum, emb_norms, ids, loss]; else [embeddings, residuals, ids, loss] -- modules/rqvae.py:118-139"""
    ctx = _Ctx()
    with ops.no_operand_cache():
        return list(ops.RqChainFunction.forward(ctx, x, mode, beta, lean, *codebooks))


@rq_chain_fwd.register_fake
def _(x, mode, beta, lean, codebooks):
    B, D = x.shape
    L = len(codebooks)
    f32 = dict(dtype=torch.float32)
    ids, loss = x.new_empty((B, L), dtype=torch.int64), x.new_empty((B,), **f32)
    if lean:
        return [x.new_empty((B, D), **f32), x.new_empty((B, L), **f32), ids, loss]
    return [x.new_empty((L, B, D), **f32), x.new_empty((L, B, D), **f32), ids, loss]


@custom_op("rqb200::rq_chain_bwd", mutates_args=(), device_types="cuda")
def rq_chain_bwd(g_a: Optional[Tensor], g_b: Optional[Tensor], g_loss: Optional[Tensor], x: Tensor, ids: Tensor, mode: int,
                 beta: float, lean: bool, codebooks: Sequence[Tensor]) -> List[Tensor]:
    L = len(codebooks)
    ctx = _Ctx((True, False, False, False, *([True] * L)))
    ctx.mode, ctx.beta, ctx.lean = mode, beta, lean
    xr = ops._rows(x)
    ctx.saved_tensors = (xr, ids, *ops._check_codebooks(codebooks, xr.shape[1]))
    gx, _m, _b, _l, *gcs = ops.RqChainFunction.backward(ctx, g_a, g_b, None, g_loss)
    return [gx, *gcs]


@rq_chain_bwd.register_fake
def _(g_a, g_b, g_loss, x, ids, mode, beta, lean, codebooks):
    return [x.new_empty(x.shape, dtype=torch.float32)] + [c.new_empty(c.shape, dtype=torch.float32) for c in codebooks]


def _rq_setup(ctx, inputs, output):
    x, mode, beta, lean, codebooks = inputs
    ctx.mode, ctx.beta, ctx.lean, ctx.L = mode, beta, lean, len(codebooks)
    ctx.save_for_backward(x, output[2], *codebooks)


def _rq_backward(ctx, grads):
    x, ids, *cbs = ctx.saved_tensors
    g_a, g_b, _g_ids, g_loss = grads
    if ctx.lean:
        g_b = None
    if g_a is None and g_b is None and g_loss is None:
        return None, None, None, None, [None] * ctx.L
    out = rq_chain_bwd(g_a, g_b, g_loss, x, ids, ctx.mode, ctx.beta, ctx.lean, list(cbs))
    return (out[0] if ctx.needs_input_grad[0] else None), None, None, None, list(out[1:])


rq_chain_fwd.register_autograd(_rq_backward, setup_context=_rq_setup)


def rq_chain(x: Tensor, mode: int, beta: float, lean: bool, codebooks: Sequence[Tensor]):
    return tuple(rq_chain_fwd(x, int(mode), float(beta), bool(lean), list(codebooks)))


# ------------------------------------------------------------------------------------------------------- Gumbel-softmax level
@custom_op("rqb200::gumbel_level_fwd", mutates_args=(), device_types="cuda")
def gumbel_level_fwd(x: Tensor, codebook: Tensor, uniform: Tensor, temperature: float, beta: float) -> List[Tensor]:
    """[emb, ids, loss, softmax weights (saved for the backward)] -- modules/quantize.py:113-136"""
    ctx = _Ctx()
    with ops.no_operand_cache():
        emb, ids, loss = ops.GumbelQuantizeFunction.forward(ctx, x, codebook, uniform, temperature, beta)
    return [emb, ids, loss, ctx.saved_tensors[2]]


@gumbel_level_fwd.register_fake
def _(x, codebook, uniform, temperature, beta):
    B, D = x.shape
    K = codebook.shape[0]
    return [x.new_empty((B, D), dtype=torch.float32), x.new_empty((B,), dtype=torch.int64),
            x.new_empty((B,), dtype=torch.float32), x.new_empty((B, K), dtype=torch.float32)]


@custom_op("rqb200::gumbel_level_bwd", mutates_args=(), device_types="cuda")
def gumbel_level_bwd(g_emb: Optional[Tensor], g_loss: Optional[Tensor], x: Tensor, codebook: Tensor, w: Tensor, emb: Tensor,
                     temperature: float, beta: float) -> List[Tensor]:
    ctx = _Ctx((True, True, False, False, False))
    ctx.temperature, ctx.beta = temperature, beta
    ctx.saved_tensors = (ops._rows(x), ops._f32c(codebook), w, emb)
    with ops.no_operand_cache():
        gx, gc, *_ = ops.GumbelQuantizeFunction.backward(ctx, g_emb, None, g_loss)
    return [gx, gc]


@gumbel_level_bwd.register_fake
def _(g_emb, g_loss, x, codebook, w, emb, temperature, beta):
    return [x.new_empty(x.shape, dtype=torch.float32), codebook.new_empty(codebook.shape, dtype=torch.float32)]


def _gumbel_setup(ctx, inputs, output):
    x, codebook, uniform, temperature, beta = inputs
    ctx.temperature, ctx.beta = temperature, beta
    ctx.save_for_backward(x, codebook, output[3], output[0])


def _gumbel_backward(ctx, grads):
    x, codebook, w, emb = ctx.saved_tensors
    g_emb, _g_ids, g_loss, _g_w = grads
    if g_emb is None and g_loss is None:
        return None, None, None, None, None
    gx, gc = gumbel_level_bwd(g_emb, g_loss, x, codebook, w, emb, ctx.temperature, ctx.beta)
    return (gx if ctx.needs_input_grad[0] else None), (gc if ctx.needs_input_grad[1] else None), None, None, None


gumbel_level_fwd.register_autograd(_gumbel_backward, setup_context=_gumbel_setup)


def gumbel_level(x: Tensor, codebook: Tensor, uniform: Tensor, temperature: float, beta: float):
    emb, ids, loss, _w = gumbel_level_fwd(x, codebook, uniform, float(temperature), float(beta))
    return emb, ids, loss


# ------------------------------------------------------------------------------------------------------------- row L2 norm
@custom_op("rqb200::l2norm_fwd", mutates_args=(), device_types="cuda")
def l2norm_fwd(x: Tensor, eps: float) -> List[Tensor]:
    """[y, row norms] over the last dim -- modules/normalize.py:6-7"""
    ctx = _Ctx()
    y = ops.L2NormFunction.forward(ctx, x, eps)
    return [_fresh(y, x), ctx.saved_tensors[1]]


@l2norm_fwd.register_fake
def _(x, eps):
    return [x.new_empty(x.shape, dtype=torch.float32), x.new_empty((x.numel() // max(x.shape[-1], 1),), dtype=torch.float32)]


@custom_op("rqb200::l2norm_bwd", mutates_args=(), device_types="cuda")
def l2norm_bwd(g: Tensor, y: Tensor, norms: Tensor, eps: float) -> Tensor:
    ctx = _Ctx()
    ctx.eps, ctx.shp = eps, y.shape
    ctx.saved_tensors = (y.reshape(-1, y.shape[-1]), norms)
    return ops.L2NormFunction.backward(ctx, g)[0]


@l2norm_bwd.register_fake
def _(g, y, norms, eps):
    return y.new_empty(y.shape, dtype=torch.float32)


def _l2_setup(ctx, inputs, output):
    ctx.eps = inputs[1]
    ctx.save_for_backward(output[0], output[1])


def _l2_backward(ctx, grads):
    y, norms = ctx.saved_tensors
    if grads[0] is None:
        return None, None
    return l2norm_bwd(grads[0].contiguous(), y, norms, ctx.eps), None


l2norm_fwd.register_autograd(_l2_backward, setup_context=_l2_setup)


def l2norm(x: Tensor, eps: float = 1e-12) -> Tensor:
    return l2norm_fwd(x, float(eps))[0]


# ----------------------------------------------------------------------------------------------- distinct id tuples (no grad)
@custom_op("rqb200::count_unique_id_tuples", mutates_args=(), device_types="cuda")
def count_unique_id_tuples(sem_ids: Tensor, codebook_size: int) -> Tensor:
    """number of distinct rows of sem_ids [B, L] as a 0-d int64 tensor -- train_rqvae.py debug statistic (rqvae.py:165-168)"""
    from .modules.rqvae import count_unique_id_tuples as impl
    return impl(sem_ids, codebook_size).reshape(()).to(torch.int64).clone()


@count_unique_id_tuples.register_fake
def _(sem_ids, codebook_size):
    return sem_ids.new_empty((), dtype=torch.int64)
