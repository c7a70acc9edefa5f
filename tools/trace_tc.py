import os
os.environ['RQB200_TC_TRACE']='1'
import sys, numpy as np, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import bench
from rq_vae_recommender_b200 import ops
x_h, cbs_h = bench.make_problem(65536)
x = torch.from_numpy(x_h).cuda(); cbs = [torch.from_numpy(c).cuda() for c in cbs_h]
state = ops.TcState(cbs)
for _ in range(300): ops.rq_tokenize_tc(x, state=state)   # long warm-up: the traced launch must run at full clocks
TL = '--timeline' in sys.argv
stats = torch.zeros(4096, dtype=torch.int32, device='cuda'); stats[3] = 1; stats[4] = 1 if TL else 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.rq_tokenize_tc(x, state=state, stats=stats); e1.record(); torch.cuda.synchronize()
s = stats.cpu().numpy()
acc = s[8:64].view(np.int64)
nb = acc[12]
names = {20:"conv wait loads landed",21:"conv convert+store",18:"epi0 merge",19:"epi0 many-pass",0:'mma wait t_empty',1:'mma wait a_full',2:'mma wait b_full',3:'mma total',4:'epi0 wait t_full',5:'epi0 scan',6:'epi0 wait pair(bar_x)',7:'epi0 merge+many+rerank',8:'epi0 total',9:'conv wait a_empty',10:'conv total',13:'epi1 wait t_full',14:'epi1 scan',15:'epi1 wait id(bar_i)',16:'epi1 other',17:'epi1 total'}
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us, blocks {nb}, rerank rows {s[0]}, cands {s[1]}, many {s[2]}")
tiles = 512
for k, n in names.items():
    print(f"{n:28s} {acc[k]/nb:12.0f} cyc/block   {acc[k]/tiles:10.0f} cyc/tile")

if TL:
    ev = s[128:128 + 2 * 256 * 4].view(np.int64).reshape(4, 256)
    names = {0: {1: 'MMA  level start (t_empty ok)', 2: 'MMA  a_full ok, chunk step', 3: 'MMA  level issued'},
             1: {1: 'CONV a_empty ok, chunk step', 2: 'CONV chunk converted+arrived'},
             2: {1: 'EPI0 t_full ok', 2: 'EPI0 scan end', 3: 'EPI0 merged', 4: 'EPI0 tmem released', 5: 'EPI0 level done (re-rank, id out)'},
             3: {1: 'EPI1 t_full ok', 2: 'EPI1 scan end', 3: 'EPI1 id received'}}
    rows = []
    for r in range(4):
        for e in ev[r]:
            e = int(e)
            if e == 0: continue
            tag, pay, clk = (e >> 56) & 0xff, (e >> 48) & 0xff, e & 0xffffffffffff
            rows.append((clk, r, tag, pay))
    rows.sort()
    t0 = rows[0][0]
    print("timeline of CTA 0 (cycles since first event; payload = tile_index*16 + level-or-step)")
    for clk, r, tag, pay in rows:
        if r in (0, 1) and tag == 2 and (pay & 15) not in (0, 11): continue      # chunk steps: first and last only
        if r == 1 and tag == 1 and (pay & 15) not in (0, 11): continue
        print(f"{clk - t0:9d}  {'    ' * r}{names[r][tag]:36s} tile {pay >> 4} {'lvl' if not (r in (0,1) and tag == 2 or r == 1) else 'step'} {pay & 15}")
