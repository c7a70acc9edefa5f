"""Event-timed split / GEMM kernels of the split-precision tensor-core path (csrc/gemm_tc.cu) at the shipped encoder shape and the
Gumbel-level GEMM shapes, next to the CUDA-core SGEMM.  gpurun -- python tools/gemm_split_time.py"""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import inputs as I
from rq_vae_recommender_b200 import ops
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
M = 65536
dims = [768, 512, 256, 128, 32]
h = torch.randn(M, 768, device="cuda") * 0.05
tot = 0
for K, N in zip(dims[:-1], dims[1:]):
    w = torch.randn(N, K, device="cuda") * 0.04
    wop = ops.SplitOperand(w)
    ts = timeit(lambda: ops.SplitOperand(h))
    aop = ops.SplitOperand(h)
    tg = timeit(lambda: ops.gemm_split(aop, wop, relu=True))
    tsg = timeit(lambda: ops.sgemm(h, w, trans_b=True, relu=True), n=3)
    fl = 2.0 * M * N * K
    print(f"M={M} K={K} N={N}: split {ts*1e3:.1f} us ({(M*K*8)/ts/1e6:.0f} GB/s), gemm {tg*1e3:.1f} us ({3*fl/tg/1e9:.0f} fp16 TFLOP/s, {fl/tg/1e9:.0f} effective), sgemm {tsg*1e3:.1f} us", flush=True)
    tot += ts + tg
    h = ops.gemm_split(aop, wop, relu=True)
print(f"sum {tot*1e3:.1f} us")
# Gumbel-level GEMMs
x = torch.randn(M, 768, device="cuda") * 0.05; cb = torch.randn(256, 768, device="cuda") * 0.05
xo, co, cto = ops.SplitOperand(x), ops.SplitOperand(cb), ops.SplitOperand(cb, transposed=True)
w = torch.rand(M, 256, device="cuda")
print(f"dist GEMM {timeit(lambda: ops.gemm_split(xo, co))*1e3:.1f} us, split W {timeit(lambda: ops.SplitOperand(w))*1e3:.1f} us, W@C {timeit(lambda: ops.gemm_split(ops.SplitOperand(w), cto))*1e3:.1f} us")
# full encoder forward + backward (dgrad + split-K wgrad on the tensor cores) at 65 536 rows
xg = (torch.randn(M, 768, device="cuda") * 0.05).requires_grad_(True)
wsg = [torch.from_numpy(w).cuda().requires_grad_(True) for w in I.mlp_weights(2, dims)]
def fb():
    y = ops.MLPFunction.apply(xg, False, *wsg)
    y.backward(torch.ones_like(y))
print(f"encoder forward + backward at {M} rows: {timeit(fb, n=5)*1e3:.1f} us", flush=True)
for (Bb, Mo, Ni) in [(65536, 512, 768), (65536, 256, 512), (4096, 512, 768)]:
    ga, hb = torch.randn(Bb, Mo, device="cuda"), torch.randn(Bb, Ni, device="cuda")
    print(f"wgrad {Mo}x{Ni} over {Bb} rows: split-K tensor cores {timeit(lambda: ops.gemm_tn(ga, hb), n=5)*1e3:.1f} us, "
          f"CUDA-core SGEMM {timeit(lambda: ops.sgemm(ga, hb, trans_a=True), n=2, ) *1e3:.1f} us", flush=True)
