"""modules/normalize.py of the reference (:6-17) on the l2norm kernels of csrc/dense.cu."""
import torch
from torch import nn
from torch import Tensor

from .. import ops


def l2norm(x, dim=-1, eps=1e-12):
    if torch.compiler.is_compiling():          # inside torch.compile: a custom-operator node instead of a graph break
        from .. import library
        if dim not in (-1, x.dim() - 1):
            return library.l2norm(x.transpose(dim, -1).contiguous(), eps).transpose(dim, -1)
        return library.l2norm(x.contiguous(), eps)
    return _l2norm_eager(x, dim, eps)


@torch.compiler.disable      # ctypes call into librqb200: opaque to Dynamo
def _l2norm_eager(x, dim=-1, eps=1e-12):
    if dim not in (-1, x.dim() - 1):
        x = x.transpose(dim, -1)
        return ops.L2NormFunction.apply(x, eps).transpose(dim, -1)
    return ops.L2NormFunction.apply(x, eps)


class L2NormalizationLayer(nn.Module):
    def __init__(self, dim=-1, eps=1e-12) -> None:
        super().__init__()
        self.dim = dim
        self.eps = eps

    def forward(self, x) -> Tensor:
        return l2norm(x, dim=self.dim, eps=self.eps)
