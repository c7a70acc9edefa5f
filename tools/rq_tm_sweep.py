"""Time the exact fused kernel for every row-tile height (RQB200_TM) at small embed dims; run each in a fresh process."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, numpy as np, torch
ROOT = %r; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
D, N = int(sys.argv[1]), int(sys.argv[2])
x, cbs = I.rq_problem(max(N, 1024), D, 256, 3, seed=9); x = x[:N]
xd = torch.from_numpy(x).cuda(); cds = [torch.from_numpy(c).cuda() for c in cbs]
for _ in range(4): ids = ops.rq_tokenize(xd, cds)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.rq_tokenize(xd, cds)
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1)/20:.4f} {int(ids.sum())}")
''' % ROOT
for D in (32, 64, 128):
    for N in (12101, 65536):
        out = []
        for tm in ("", "8", "4", "2", "1"):
            env = dict(os.environ); env.pop("RQB200_TM", None)
            if tm: env["RQB200_TM"] = tm
            r = subprocess.run([sys.executable, "-c", CHILD, str(D), str(N)], env=env, capture_output=True, text=True)
            ms, chk = r.stdout.split() if r.returncode == 0 else ("nan", r.stderr[-80:])
            out.append(f"TM={tm or 'auto'}: {ms} ms")
            out.append(f"[{chk}]") if tm == "" else None
        print(f"D={D} N={N}: " + "  ".join(o for o in out if o))
