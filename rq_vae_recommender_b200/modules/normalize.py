"""modules/normalize.py of the reference (:6-17) on the l2norm kernels of csrc/dense.cu."""
import torch
from torch import nn
from torch import Tensor

from .. import ops


@torch.compiler.disable      # ctypes call into librqb200: opaque to Dynamo
def l2norm(x, dim=-1, eps=1e-12):
    if dim not in (-1, x.dim() - 1):
        x = x.transpose(dim, -1)
        return ops.L2NormFunction.apply(x, eps).transpose(dim, -1)
    return ops.L2NormFunction.apply(x, eps)


class L2NormalizationLayer(nn.Module):
    def __init__(self, dim=-1, eps=1e-12) -> None:
        super().__init__()
        self.dim = dim
        self.eps = eps

    @torch.compiler.disable      # ctypes call into librqb200: opaque to Dynamo
    def forward(self, x) -> Tensor:
        return l2norm(x, dim=self.dim, eps=self.eps)
