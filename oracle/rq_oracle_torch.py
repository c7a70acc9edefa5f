"""CPU port of the reference tokenisation path in torch CPU ops  --  TEST / BASELINE INFRASTRUCTURE ONLY.

Same role and rules as oracle/rq_oracle.py (only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs may import it).  It exists because the reference IS torch code: timing a numpy port would
under-state the reference's CPU speed (numpy elementwise ops are single threaded, ATen's are not).  Every statement
below is the reference's own expression, op for op, so its runtime on the host cores is what the reference's eval
path costs there.  Checked against the numpy oracle (and through it the golden fixtures) in tests/test_oracle_golden.py.
"""
import torch


@torch.no_grad()
def quantize_eval(x: torch.Tensor, codebook: torch.Tensor):
    """modules/quantize.py:113-128,159-160 (eval branch, L2): returns (ids, emb_out)."""
    dist = (
        (x**2).sum(axis=1, keepdim=True)
        + (codebook.T**2).sum(axis=0, keepdim=True)
        - 2 * x @ codebook.T
    )
    _, ids = (dist.detach()).min(axis=1)
    return ids, codebook[ids]


@torch.no_grad()
def rq_tokenize(x: torch.Tensor, codebooks):
    """modules/rqvae.py:125-132 in eval mode, sem_ids only (what semids.py:125 consumes)."""
    res = x
    sem_ids = []
    for cb in codebooks:
        ids, emb = quantize_eval(res, cb)
        res = res - emb
        sem_ids.append(ids)
    return torch.stack(sem_ids, dim=1)
