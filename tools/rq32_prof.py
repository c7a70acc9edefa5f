"""A few launches of the exact fused kernel at the shipped quantiser shape (D=32, K=256, L=3) for ncu."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
x, cbs = I.rq_problem(65536, 32, 256, 3, seed=9)
xd = torch.from_numpy(x).cuda(); cds = [torch.from_numpy(c).cuda() for c in cbs]
for _ in range(4): ops.rq_tokenize(xd, cds)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.rq_tokenize(xd, cds)
e1.record(); torch.cuda.synchronize()
print(f"rq_tokenize 65536x32 K=256 L=3: {e0.elapsed_time(e1)/10:.3f} ms per call (3 launches: transpose, norms, fused)")
