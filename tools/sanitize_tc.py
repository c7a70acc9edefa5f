import sys, numpy as np, torch, os
ROOT = '/root/repo' if os.path.isdir('/root/repo') else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
for (B, D, L) in [(1, 768, 3), (129, 768, 3), (300, 64, 1), (513, 256, 4), (385, 768, 3)]:
    x, cbs = I.rq_problem(max(B, 1024), D, 256, L, seed=B + D); x = x[:B]
    st = torch.zeros(64, dtype=torch.int32, device='cuda')
    ids = ops.rq_tokenize_tc(dev(x), [dev(c) for c in cbs], stats=st)
    ref = ops.rq_tokenize(dev(x), [dev(c) for c in cbs])
    print(B, D, L, 'rows equal to exact kernel', float((ids == ref).all(dim=1).float().mean()), flush=True)
# misaligned x (scalar-load instantiation)
x, cbs = I.rq_problem(1024, 128, 256, 2, seed=9)
buf = torch.zeros(200 * 129 + 1, dtype=torch.float32, device='cuda')
xv = buf[1:1 + 200 * 129].view(200, 129)[:, :128]
xv.copy_(dev(x[:200]))
ids = ops.rq_tokenize_tc(xv, [dev(c) for c in cbs])
torch.cuda.synchronize(); print("tc sanitize pass done")
