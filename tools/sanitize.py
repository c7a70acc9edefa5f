"""Small-shape pass over every kernel family for compute-sanitizer memcheck."""
import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
from rq_vae_recommender_b200.init.kmeans import Kmeans
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
# fused chain fwd (all modes, ragged shapes) + bwd
for (B, D, K, L) in [(37, 20, 5, 3), (130, 64, 256, 3), (65, 768, 256, 3)]:
    x, cbs = I.rq_problem(max(B, K), D, K, L, seed=B); x = x[:B]
    for mode in (0, 2, 3):
        xt = dev(x).requires_grad_(True); cts = [dev(c).requires_grad_(True) for c in cbs]
        a, b, ids, loss = ops.RqChainFunction.apply(xt, mode, 0.25, False, *cts)
        (a.sum() + b.sum() + loss.sum()).backward()
# tc tokeniser: partial last tile, several tiles, L=1 and L=4
for (B, D, L) in [(1, 768, 3), (129, 768, 3), (300, 64, 1), (513, 256, 4)]:
    x, cbs = I.rq_problem(max(B, 1024), D, 256, L, seed=B + D); x = x[:B]
    st = torch.zeros(64, dtype=torch.int32, device='cuda')
    ops.rq_tokenize_tc(dev(x), [dev(c) for c in cbs], stats=st)
# gumbel level fwd/bwd, mlp fwd/bwd, kmeans, histogram
x, cbs = I.rq_problem(300, 32, 256, 1, seed=3)
xt = dev(x).requires_grad_(True); ct = dev(cbs[0]).requires_grad_(True)
e, ids, loss = ops.GumbelQuantizeFunction.apply(xt, ct, dev(I.rand(4, 300, 256)), 0.2, 0.25); (e.sum() + loss.sum()).backward()
ws = [dev(w).requires_grad_(True) for w in I.mlp_weights(5, [70, 33, 17])]
xm = dev(I.randn(6, 45, 70)).requires_grad_(True)
ops.MLPFunction.apply(xm, True, *ws).sum().backward()
np.random.seed(0); torch.manual_seed(0)
Kmeans(k=16, max_iters=3).run(dev(I.randn(7, 500, 12)))
ids3 = torch.randint(0, 256, (1000, 3), device='cuda')
ops.sid_histogram(ids3, 256)
rank, st = ops.sid_dedup_rank(ids3, 256)
ops.sid_gather(torch.cat([ids3, rank.unsqueeze(1)], 1), torch.randint(0, 1000, (7, 20), device='cuda'), torch.rand(7, 20, device='cuda') > 0.3)
# tc tokeniser at both tile shapes: several tiles per CTA pair, ragged tail
x, cbs = I.rq_problem(30000, 768, 256, 3, seed=12)
ops.rq_tokenize_tc(dev(x[:777]), [dev(c) for c in cbs]); ops.rq_tokenize_tc(dev(x), [dev(c) for c in cbs])
# split-precision GEMM: ragged M / N / K, transposed operand, mask; MLP forward + dgrad on it; Gumbel level on it
a = dev(I.randn(20, 700, 100)); b = dev(I.randn(21, 200, 100)); bt = dev(I.randn(22, 100, 130))
ops.gemm_split(a, b, relu=True); ops.gemm_split(a, ops.SplitOperand(bt, transposed=True), mask=dev(I.randn(23, 700, 130)))
ops.gemm_tn(dev(I.randn(28, 700, 130)), dev(I.randn(29, 700, 75)))          # transposed splits + split-K + reduce
ws = [dev(w).requires_grad_(True) for w in I.mlp_weights(24, [72, 40, 24])]
xm = dev(I.randn(25, 600, 72)).requires_grad_(True)
ops.MLPFunction.apply(xm, True, *ws).sum().backward()
x, cbs = I.rq_problem(640, 64, 256, 1, seed=26)
xt = dev(x).requires_grad_(True); ct = dev(cbs[0]).requires_grad_(True)
e, ids, loss = ops.GumbelQuantizeFunction.apply(xt, ct, dev(I.rand(27, 640, 256)), 0.2, 0.25); (e.sum() + loss.sum()).backward()
# prefix index + beam selection (ragged batch, out-of-range ids)
corp = torch.randint(0, 16, (333, 3), device='cuda'); pidx = ops.SidPrefixIndex(corp, 16)
pidx.check(torch.randint(-2, 18, (1001, 2), device='cuda'))
g1, p1, _ = pidx.beam_select(torch.randint(0, 16, (5, 16), device='cuda'), torch.randn(5, 16, device='cuda'), None, None, 3)
pidx.beam_select(torch.randint(0, 16, (15, 16), device='cuda'), torch.randn(15, 16, device='cuda'), g1, p1, 3)
torch.cuda.synchronize(); print("sanitize pass done")
