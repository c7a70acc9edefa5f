"""A/B timing of the TC tokeniser's run-time switches on ONE box (boxes differ by several percent).
Each configuration runs in its own process because the library reads the switches once.
usage: python tools/tc_ab.py            -> spawns the 4 combinations of RQB200_TC_ROT x RQB200_TC_PREFETCH
       python tools/tc_ab.py --one      -> times the current environment (used by the spawner)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    import bench
    from rq_vae_recommender_b200 import ops
    x_h, cbs_h = bench.make_problem(65536)
    x = torch.from_numpy(x_h).cuda()
    cbs = [torch.from_numpy(c).cuda() for c in cbs_h]
    state = ops.TcState(cbs)
    for _ in range(10):
        ids = ops.rq_tokenize_tc(x, state=state)
    best = 1e9
    tot = 0.0
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
    print(f"ROT={os.environ.get('RQB200_TC_ROT','0')} PREFETCH={os.environ.get('RQB200_TC_PREFETCH','0')}: "
          f"best {best*1e3:.1f} us  mean {tot/5*1e3:.1f} us  checksum {int(ids.sum())}", flush=True)


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for rot in "01":
            for pf in "01":
                env = dict(os.environ, RQB200_TC_ROT=rot, RQB200_TC_PREFETCH=pf)
                subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, check=True)
