"""One pass over the kernel families that are not the tokeniser, at representative shapes, as an ncu target:
    ncu --set full --clock-control none -k regex:'gs_gemm|gs_split_rows|rq_fused|sgemm_kernel|rq_bwd|gumbel' ... python tools/kernels_prof.py
Shapes: the shipped encoder at 65 536 rows (split GEMMs), the fused STE chain forward + backward at 4096 x 32 (C1-like) and
65 536 x 32, a Gumbel-softmax level forward + backward at 16 384 x 768 x 256 (C4-like), an SGEMM below the tensor-core threshold."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    # encoder forward + dgrad on the split GEMM
    x = (torch.randn(65536, 768, device="cuda") * 0.05).requires_grad_(True)
    ws = [dev(w).requires_grad_(True) for w in I.mlp_weights(2, [768, 512, 256, 128, 32])]
    z = ops.MLPFunction.apply(x, False, *ws)
    z.sum().backward()
    # fused chain (STE) forward + backward at the shipped quantiser width
    zz = z.detach().requires_grad_(True)
    cbs = [(torch.randn(256, 32, device="cuda") * float(z.std()) * 0.6 ** l).requires_grad_(True) for l in range(3)]
    a, b, ids, loss = ops.RqChainFunction.apply(zz, ops.MODE_STE, 0.25, True, *cbs)
    (a.sum() + loss.sum()).backward()
    # Gumbel-softmax level forward + backward (C4 shape per level, a quarter of the rows)
    xg = (torch.randn(16384, 768, device="cuda") * 0.05).requires_grad_(True)
    cg = (torch.randn(256, 768, device="cuda") * 0.05).requires_grad_(True)
    e, gi, gl = ops.GumbelQuantizeFunction.apply(xg, cg, torch.rand(16384, 256, device="cuda"), 0.2, 0.25)
    (e.sum() + gl.sum()).backward()
    # corpus-side kernels: dedup rank + statistics, sequence gather, prefix index + beam selection
    ids3 = torch.randint(0, 256, (84000, 3), device="cuda")
    rank, st = ops.sid_dedup_rank(ids3, 256)
    table = torch.cat([ids3, rank.clamp_max(255).unsqueeze(1)], 1)
    ops.sid_gather(table, torch.randint(0, 84000, (256, 20), device="cuda"), torch.rand(256, 20, device="cuda") > 0.2)
    pidx = ops.SidPrefixIndex(table, 256)
    pidx.check(table[torch.randint(0, 84000, (163840,), device="cuda")])
    g1, p1, _ = pidx.beam_select(torch.randint(0, 256, (256, 64), device="cuda"), torch.randn(256, 64, device="cuda"), None, None, 10)
    pidx.beam_select(torch.randint(0, 256, (2560, 64), device="cuda"), torch.randn(2560, 64, device="cuda"), g1, p1, 10)
    # small-batch SGEMM (the reference's training batch sizes)
    ops.sgemm(torch.randn(256, 768, device="cuda"), ws[0].detach(), trans_b=True, relu=True)
torch.cuda.synchronize(); print("done")
