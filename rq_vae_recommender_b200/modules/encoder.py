"""modules/encoder.py of the reference (:7-38): bias-free Linear+ReLU stack, optional final L2 norm.

The module tree (``self.mlp`` Sequential with Linear at indices 0,2,4,6) is kept so state-dict keys match the
shipped checkpoints; forward() bypasses it and runs the whole stack as one autograd node on the fp32 GEMM kernel
with the ReLU fused in its epilogue (ops.MLPFunction)."""
from typing import List

from torch import nn
from torch import Tensor

from .. import ops
from .normalize import L2NormalizationLayer


class MLP(nn.Module):
    def __init__(self, input_dim: int, hidden_dims: List[int], out_dim: int, dropout: float = 0.0,
                 normalize: bool = False) -> None:
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dims = hidden_dims
        self.out_dim = out_dim
        self.dropout = dropout
        self.normalize = normalize

        dims = [self.input_dim] + list(self.hidden_dims) + [self.out_dim]
        self.mlp = nn.Sequential()
        for i, (in_d, out_d) in enumerate(zip(dims[:-1], dims[1:])):
            self.mlp.append(nn.Linear(in_d, out_d, bias=False))
            if i != len(dims) - 2:
                self.mlp.append(nn.ReLU())
                if dropout != 0:
                    self.mlp.append(nn.Dropout(dropout))
        self.mlp.append(L2NormalizationLayer() if normalize else nn.Identity())

    def forward(self, x: Tensor) -> Tensor:
        assert x.shape[-1] == self.input_dim, f"Invalid input dim: Expected {self.input_dim}, found {x.shape[-1]}"
        if self.dropout != 0 and self.training:
            raise NotImplementedError("MLP dropout > 0 in training is not built (no reference caller sets it)")
        weights = [m.weight for m in self.mlp if isinstance(m, nn.Linear)]
        lead = x.shape[:-1]
        y = ops.MLPFunction.apply(x.reshape(-1, self.input_dim), bool(getattr(self, "normalize", False)) or
                                  isinstance(self.mlp[-1], L2NormalizationLayer), *weights)
        return y.reshape(*lead, self.out_dim)
