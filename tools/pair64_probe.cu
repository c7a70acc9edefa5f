// Bring-up probe for rq_tc64_kernel (csrc/rq_tc64.cu), the 64-rows-per-CTA tokeniser written at the end of round 1 without
// GPU access.  It answers, on ONE cluster of two CTAs and in a few milliseconds, the three questions that kernel rests on:
//   1. layout   tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA), N = 256: where do the accumulators land in TMEM?
//               The kernel assumes cute's "2x2" atom: (m, n) -> lane m + 64 (n / 128), column n % 128.  The probe dumps all
//               128 lanes x 256 columns of both CTAs and reports which of the candidate layouts reproduces an exact GEMM.
//   2. x path   a 2-D tensor map over fp32 x with a 64 x 64 box (rows past the end zero-filled), converted in the kernel to
//               the fp16 K-major SWIZZLE_128B image with the SAME index functions the kernel uses (csrc/tc64_layout.cuh)
//   3. rate     cycles per tcgen05.mma for M = 128 vs M = 256 pair instructions (N = 256, K = 16): the design needs M = 128
//               to run at the full per-SM rate (64 cycles per instruction; 128 would mean half rate -> drop the design)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/pair64_probe tools/pair64_probe.cu
// run:   tools/bin/pair64_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../rq_vae_recommender_b200/csrc/tc64_layout.cuh"

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, int tag, int* dbg) {
  const long long t0 = clock64();
  while (!mbar_try(b, parity)) {
    if (clock64() - t0 > 1000000000LL) { if (dbg) dbg[8 + tag] = 1 + (int)cta_rank(); __threadfence_system(); __trap(); }
  }
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"(tm), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2d_cg2(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"(tm), "r"(mbar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t a) {   // K-major SWIZZLE_128B, SBO 1024 B (same as csrc/tc_common.cuh)
  return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(s32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st32_zero(uint32_t taddr) {   // clear 32 columns of this thread's lane
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n\t"
      "tcgen05.wait::st.sync.aligned;"
      ::"r"(taddr), "r"(0x7fc00000u) : "memory");   // NaN pattern: untouched cells stay recognisable
}

struct Smem {
  alignas(1024) uint8_t A[16384];     // 8 KB used by the M = 128 run (64 rows), 16 KB by the M = 256 timing run
  alignas(1024) uint8_t B[16384];
  alignas(1024) float X[64 * 64];     // fp32 staging box
  uint64_t xfull;   // local: the fp32 box landed
  uint64_t afull;   // leader: both CTAs converted their A (count 2, remote arrive from the peer)
  uint64_t bfull;   // leader: both B halves landed (32 KB of tensor-TMA bytes from both CTAs)
  uint64_t done;    // both: the MMAs completed (multicast commit)
  uint64_t tdone;   // both: timing loop completed
  uint32_t tmem;
};

// x: [xrows][64] fp32 (xrows may be < 128: rows past the end must read as zero); B images as in pair_probe.cu
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair64_probe(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapB, float* Draw, long long* cyc,
             int iters, int* dbg) {
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  Smem& sm = *reinterpret_cast<Smem*>(dyn_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cta_rank();
  if (tid == 0) {
    mbar_init(&sm.xfull, 1); mbar_init(&sm.afull, 2); mbar_init(&sm.bfull, 1); mbar_init(&sm.done, 1); mbar_init(&sm.tdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&sm.tmem)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&sm.tmem);
  // poison all 256 columns so that cells the MMA does not write are recognisable in the dump
  for (int c = 0; c < 256; c += 32) st32_zero(tbase + ((uint32_t)(warp * 32) << 16) + c);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    mbar_expect(&sm.xfull, 64 * 64 * 4);
    tma2d(sm.X, &tmapX, 0, (int)rank * 64, &sm.xfull);                 // this CTA's 64 rows of x
    if (rank == 0) mbar_expect(&sm.bfull, 32768);
    tma2d_cg2(sm.B, &tmapB, 0, (int)rank * 128, mapa(s32(&sm.bfull), 0));
  }
  mbar_wait(&sm.xfull, 0, 0, dbg);
  // the kernel's converter: 4 warps x 16 rows, step j covers rows 16 w + 2 j + (lane >> 4), float4 column lane & 15
  for (int j = 0; j < 8; ++j) {
    const int r = 16 * warp + 2 * j + (lane >> 4), q = lane & 15;
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(sm.X) + tc64_stage_offset(r, q));
    const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s32(sm.A) + tc64_a_offset(r, q)),
                 "r"(*reinterpret_cast<const uint32_t*>(&h0)), "r"(*reinterpret_cast<const uint32_t*>(&h1)) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) mbar_arrive_cluster(mapa(s32(&sm.afull), 0));

  if (rank == 0 && warp == 0) {
    if (lane == 0) {
      mbar_wait(&sm.afull, 0, 1, dbg);
      mbar_wait(&sm.bfull, 0, 2, dbg);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // M = 128, N = 256
      const uint64_t ad = smem_desc(s32(sm.A)), bd = smem_desc(s32(sm.B));
      for (int j = 0; j < 4; ++j) mma2(tbase, ad + 2 * j, bd + 2 * j, idesc, j != 0);
      commit2(&sm.done, 3);
    }
    __syncwarp();
  }
  mbar_wait(&sm.done, 0, 3, dbg);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // raw dump: Draw[rank][lane 0..127][column 0..255]
  for (int c = 0; c < 256; c += 32) {
    uint32_t r[32];
    ld32(tbase + ((uint32_t)(warp * 32) << 16) + c, r);
    for (int e = 0; e < 32; ++e) Draw[((size_t)rank * 128 + warp * 32 + lane) * 256 + c + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  // ---- rate: `iters` x 4 instructions into the same accumulators, M = 128 then M = 256 (the A buffer holds 128 rows of
  // whatever is there: only the timing matters).  cyc[0] / cyc[1] = cycles from first issue to completion.
  for (int shape = 0; shape < 2; ++shape) {
    if (rank == 0 && warp == 0) {
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int M = shape == 0 ? 128 : 256;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t ad = smem_desc(s32(sm.A)), bd = smem_desc(s32(sm.B));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i)
          for (int j = 0; j < 4; ++j) mma2(tbase, ad + 2 * j, bd + 2 * j, idesc, 1);
        commit2(&sm.tdone, 3);
        mbar_wait(&sm.tdone, shape & 1, 4, dbg);
        cyc[shape] = clock64() - t0;
      }
      __syncwarp();
    }
    if (!(rank == 0 && warp == 0 && lane == 0)) mbar_wait(&sm.tdone, shape & 1, 5, dbg);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int xrows = argc > 1 ? atoi(argv[1]) : 100;      // < 128: rows [xrows, 128) must come back as zero from the tensor map
  const int iters = argc > 2 ? atoi(argv[2]) : 4096;
  // integer data in [-4, 4]: every product and sum is exact in fp16 / fp32
  std::vector<float> X((size_t)xrows * 64), Bm(256 * 64);
  srand(11);
  for (auto& v : X) v = (float)(rand() % 9 - 4);
  for (auto& v : Bm) v = (float)(rand() % 9 - 4);
  std::vector<__half> bimg(256 * 64);
  for (int r = 0; r < 256; ++r)
    for (int k = 0; k < 64; ++k) {
      const int blk = r / 128, rr = r % 128;
      bimg[blk * 8192 + rr * 64 + (((k >> 3) ^ (rr & 7)) * 8) + (k & 7)] = __float2half(Bm[r * 64 + k]);
    }
  float *dx, *dD; uint8_t* db; int* ddbg; long long* dcyc;
  CK(cudaMalloc(&dx, X.size() * 4)); CK(cudaMalloc(&db, 32768)); CK(cudaMalloc(&dD, 2 * 128 * 256 * 4));
  CK(cudaMalloc(&ddbg, 64 * 4)); CK(cudaMalloc(&dcyc, 16));
  CK(cudaMemcpy(dx, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bimg.data(), 32768, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 2 * 128 * 256 * 4)); CK(cudaMemset(ddbg, 0, 64 * 4)); CK(cudaMemset(dcyc, 0, 16));

  CUtensorMap tmB, tmX; memset(&tmB, 0, sizeof(tmB)); memset(&tmX, 0, sizeof(tmX));
  {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    const cuuint32_t estr[2] = {1, 1};
    {
      const cuuint64_t gdim[2] = {64, 256}; const cuuint64_t gstr[1] = {128}; const cuuint32_t box[2] = {64, 128};
      const CUresult r = ((EncodeFn)fn)(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode B failed: %d\n", (int)r); return 2; }
    }
    {
      const cuuint64_t gdim[2] = {64, (cuuint64_t)xrows}; const cuuint64_t gstr[1] = {256}; const cuuint32_t box[2] = {64, 64};
      const CUresult r = ((EncodeFn)fn)(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode X failed: %d\n", (int)r); return 2; }
    }
  }
  CK(cudaFuncSetAttribute(pair64_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  pair64_probe<<<2, 128, sizeof(Smem)>>>(tmX, tmB, dD, dcyc, iters, ddbg);
  const cudaError_t le = cudaGetLastError();
  const cudaError_t se = cudaDeviceSynchronize();
  int dbg[64]; cudaMemcpy(dbg, ddbg, sizeof(dbg), cudaMemcpyDeviceToHost);
  printf("launch=%s sync=%s timeouts[xfull,afull,bfull,done,tdone,tdone']=%d %d %d %d %d %d\n", cudaGetErrorString(le),
         cudaGetErrorString(se), dbg[8], dbg[9], dbg[10], dbg[11], dbg[12], dbg[13]);
  if (se != cudaSuccess) return 1;
  std::vector<float> Dh(2 * 128 * 256);
  CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
  long long cyc[2]; CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
  auto ref = [&](int i, int n) {   // row i of the 128-row pair tile (zero past xrows) . code n
    float s = 0.f;
    if (i < xrows) for (int k = 0; k < 64; ++k) s += X[(size_t)i * 64 + k] * Bm[n * 64 + k];
    return s;
  };
  // candidate layouts for element (m, n) of CTA r's 64 x 256 tile -> (lane, column)
  struct Cand { const char* name; int (*lane)(int, int); int (*col)(int, int); };
  const Cand cands[] = {
    {"2x2 (kernel's assumption): lane m + 64 (n / 128), col n % 128", [](int m, int n) { return tc64_tmem_lane(m, n); }, [](int, int n) { return tc64_tmem_col(n); }},
    {"rows in lanes 0..63, col n", [](int m, int) { return m; }, [](int, int n) { return n; }},
    {"M=64 single-CTA style: lane (m % 16) + 32 (m / 16), col n", [](int m, int) { return (m % 16) + 32 * (m / 16); }, [](int, int n) { return n; }},
    {"2x2 with the column halves swapped: lane m + 64 (1 - n / 128)", [](int m, int n) { return m + 64 * (1 - (n >> 7)); }, [](int, int n) { return n & 127; }},
  };
  int winner = -1;
  for (int c = 0; c < 4; ++c) {
    long bad = 0;
    for (int r = 0; r < 2; ++r)
      for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 256; ++n) {
          const float got = Dh[((size_t)r * 128 + cands[c].lane(m, n)) * 256 + cands[c].col(m, n)];
          if (!(got == ref(64 * r + m, n))) ++bad;
        }
    printf("layout %-70s mismatches %ld / 32768\n", cands[c].name, bad);
    if (bad == 0 && winner < 0) winner = c;
  }
  printf("cycles for %d x 4 instructions (N=256, K=16): M=128 pair %lld (%.1f / instr), M=256 pair %lld (%.1f / instr)\n", iters,
         cyc[0], (double)cyc[0] / (4.0 * iters), cyc[1], (double)cyc[1] / (4.0 * iters));
  printf("expected if M=128 runs at the full per-SM rate: ~64 / instr vs ~128 / instr for M=256\n");
  printf("%s\n", winner == 0 ? "PAIR64 PROBE OK" : (winner > 0 ? "PAIR64 PROBE: OTHER LAYOUT (see above)" : "PAIR64 PROBE MISMATCH"));
  if (winner != 0) {
    printf("cta0 lane 0 cols 0..3: %g %g %g %g | lane 64 cols 0..3: %g %g %g %g | lane 0 cols 128..131: %g %g %g %g ; ref(0,0..3) %g %g %g %g ref(0,128) %g\n",
           Dh[0], Dh[1], Dh[2], Dh[3], Dh[64 * 256], Dh[64 * 256 + 1], Dh[64 * 256 + 2], Dh[64 * 256 + 3], Dh[128], Dh[129], Dh[130], Dh[131],
           ref(0, 0), ref(0, 1), ref(0, 2), ref(0, 3), ref(0, 128));
  }
  return winner == 0 ? 0 : 1;
}
