"""A/B timing of the TC tokeniser's run-time switches on ONE box (boxes differ by several percent).
Each configuration runs in its own process because the library reads the switches once.
usage: python tools/tc_ab.py CONFIG [CONFIG ...]   with CONFIG = comma-separated NAME=VALUE pairs ("" = defaults), e.g.
       python tools/tc_ab.py "" RQB200_TC_PAIR=1 RQB200_TC_ROT=1,RQB200_TC_PREFETCH=1
       python tools/tc_ab.py --one    times the current environment (used by the spawner)
switches: RQB200_TC_PAIR (CTA-pair kernel), RQB200_TC_ROT (per-CTA k-chunk rotation), RQB200_TC_PREFETCH (L2 prefetch of x)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import bench
    from rq_vae_recommender_b200 import ops
    n = int(os.environ.get("TC_AB_ROWS", "65536"))
    x_h, cbs_h = bench.make_problem(n)
    x = torch.from_numpy(x_h).cuda()
    cbs = [torch.from_numpy(c).cuda() for c in cbs_h]
    state = ops.TcState(cbs)
    ref = ops.rq_tokenize(x[:4096], cbs)            # exact CUDA-core kernel on a slice: sanity check of the variant under test
    for _ in range(10):
        ids = ops.rq_tokenize_tc(x, state=state)
    agree = float((ids[:4096] == ref).all(dim=1).float().mean())
    best, tot = 1e9, 0.0
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ids = ops.rq_tokenize_tc(x, state=state)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        best = min(best, ms)
        tot += ms
    cfg = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("RQB200_TC_"))
    print(f"[{cfg or 'defaults'}] rows {n}: best {best*1e3:.1f} us  mean {tot/5*1e3:.1f} us  checksum {int(ids.sum())} "
          f"rows equal to the exact kernel (first 4096): {agree:.4f}", flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for cfg in (sys.argv[1:] or [""]):
            env = dict(os.environ)
            for kv in filter(None, cfg.split(",")):
                k, v = kv.split("=")
                env[k] = v
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env)
            if r.returncode:
                print(f"[{cfg}] FAILED rc={r.returncode}", flush=True)
