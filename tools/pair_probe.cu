// Standalone bring-up probe for the CTA-pair (cta_group::2) tensor-core path planned for the tokeniser.
// One 256 x 256 x 64 fp16 GEMM on ONE cluster of two CTAs, checked exactly against a CPU reference (small integer data).
// It isolates every mechanism the pair kernel needs, with guarded waits so a protocol error traps instead of hanging:
//   * cluster launch (__cluster_dims__), cluster-scope barrier init / sync, mapa + remote mbarrier.arrive
//   * tcgen05.alloc / dealloc .cta_group::2 issued by warp 0 of BOTH CTAs
//   * tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA; N = 256: 128 B columns held by each CTA) issued by the leader
//   * tcgen05.commit .cta_group::2 .multicast::cluster (completion delivered to both CTAs)
//   * mode 1: B loaded by tensor-map TMA with .cta_group::2, the peer's bytes signalling the LEADER's mbarrier
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scratch/pair_probe tools/pair_probe.cu
// run:   scratch/pair_probe [mode 0|1]
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // local or remote, cluster-scope release
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma2d_cg2(void* dst, const CUtensorMap* tm, int c0, int c1, uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"(tm), "r"(mbar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t a) {   // K-major SWIZZLE_128B, SBO 1024 B (same as csrc/rq_tc.cu)
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

struct Smem {
  alignas(1024) uint8_t A[16384];
  alignas(1024) uint8_t B[16384];
  uint64_t full;    // local: own A (and own B in mode 0) landed
  uint64_t afull;   // leader: both CTAs report their A resident (count 2, remote arrive from the peer)
  uint64_t bfull;   // leader: both B halves landed (mode 1: 32 KB of tensor-TMA bytes from both CTAs)
  uint64_t done;    // both: the MMAs completed (multicast commit)
  uint32_t tmem;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe(const uint8_t* a_img, const uint8_t* b_img, float* D, const __grid_constant__ CUtensorMap tmapB, int mode, int* dbg) {
  __shared__ Smem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cta_rank();
  if (tid == 0) {
    mbar_init(&sm.full, 1); mbar_init(&sm.afull, 2); mbar_init(&sm.bfull, 1); mbar_init(&sm.done, 1);
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
  if (tid == 0) { dbg[rank] = (int)tbase; dbg[2 + rank] = (int)s32(&sm.full); dbg[4 + rank] = (int)mapa(s32(&sm.full), 0); }

  if (tid == 0) {
    if (mode == 0) {
      mbar_expect(&sm.full, 32768);
      bulk_g2s(sm.A, a_img + rank * 16384, 16384, &sm.full);
      bulk_g2s(sm.B, b_img + rank * 16384, 16384, &sm.full);
    } else {
      mbar_expect(&sm.full, 16384);
      bulk_g2s(sm.A, a_img + rank * 16384, 16384, &sm.full);
      if (rank == 0) mbar_expect(&sm.bfull, 32768);
      tma2d_cg2(sm.B, &tmapB, 0, (int)rank * 128, mapa(s32(&sm.bfull), 0));
    }
    mbar_wait(&sm.full, 0, 0, dbg);
    mbar_arrive_cluster(mapa(s32(&sm.afull), 0));         // tell the leader this CTA's operands are resident
  }
  if (rank == 0 && warp == 0) {
    if (lane == 0) {
      mbar_wait(&sm.afull, 0, 1, dbg);
      if (mode == 1) mbar_wait(&sm.bfull, 0, 2, dbg);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t ad = smem_desc(s32(sm.A)), bd = smem_desc(s32(sm.B));
      for (int j = 0; j < 4; ++j) mma2(tbase, ad + 2 * j, bd + 2 * j, idesc, j != 0);
      commit2(&sm.done, 3);
    }
    __syncwarp();
  }
  mbar_wait(&sm.done, 0, 3, dbg);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = (int)rank * 128 + warp * 32 + lane;
  for (int c = 0; c < 256; c += 32) {
    uint32_t r[32];
    ld32(tbase + ((uint32_t)(warp * 32) << 16) + c, r);
    for (int e = 0; e < 32; ++e) D[row * 256 + c + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256u) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  // integer data in [-4, 4]: every product and sum is exact in fp16 / fp32
  std::vector<float> A(256 * 64), B(256 * 64);
  srand(7);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : B) v = (float)(rand() % 9 - 4);
  auto image = [](const std::vector<float>& M) {          // two 128-row K-major SWIZZLE_128B images of 16 KB
    std::vector<__half> img(256 * 64);
    for (int r = 0; r < 256; ++r)
      for (int k = 0; k < 64; ++k) {
        const int blk = r / 128, rr = r % 128;
        img[blk * 8192 + rr * 64 + (((k >> 3) ^ (rr & 7)) * 8) + (k & 7)] = __float2half(M[r * 64 + k]);
      }
    return img;
  };
  const auto ai = image(A), bi = image(B);
  uint8_t *da, *db; float* dD; int* ddbg;
  CK(cudaMalloc(&da, 32768)); CK(cudaMalloc(&db, 32768)); CK(cudaMalloc(&dD, 256 * 256 * 4)); CK(cudaMalloc(&ddbg, 64 * 4));
  CK(cudaMemcpy(da, ai.data(), 32768, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), 32768, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 256 * 256 * 4)); CK(cudaMemset(ddbg, 0, 64 * 4));

  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    const cuuint64_t gdim[2] = {64, 256};            // innermost first: 64 halves (128 B) per image row, 256 image rows
    const cuuint64_t gstr[1] = {128};                // byte stride of dimension 1
    const cuuint32_t box[2] = {64, 128};             // one 16 KB image
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 2; }
  }
  pair_probe<<<2, 128>>>(da, db, dD, tm, mode, ddbg);
  const cudaError_t le = cudaGetLastError();
  const cudaError_t se = cudaDeviceSynchronize();
  int dbg[64]; cudaMemcpy(dbg, ddbg, sizeof(dbg), cudaMemcpyDeviceToHost);
  printf("mode %d launch=%s sync=%s tmem base cta0=0x%x cta1=0x%x smem(full) cta0=0x%x cta1=0x%x mapa->0: 0x%x 0x%x timeouts[full,afull,bfull,done]=%d %d %d %d\n",
         mode, cudaGetErrorString(le), cudaGetErrorString(se), dbg[0], dbg[1], dbg[2], dbg[3], dbg[4], dbg[5], dbg[8], dbg[9], dbg[10], dbg[11]);
  if (se != cudaSuccess) return 1;
  std::vector<float> Dh(256 * 256);
  CK(cudaMemcpy(Dh.data(), dD, 256 * 256 * 4, cudaMemcpyDeviceToHost));
  double worst[2][2] = {{0, 0}, {0, 0}};
  for (int i = 0; i < 256; ++i)
    for (int j = 0; j < 256; ++j) {
      float ref = 0.f;
      for (int k = 0; k < 64; ++k) ref += A[i * 64 + k] * B[j * 64 + k];
      const double e = fabs((double)Dh[i * 256 + j] - ref);
      if (!(e <= worst[i / 128][j / 128])) worst[i / 128][j / 128] = (e == e) ? e : 1e30;
    }
  printf("max |D - ref| by quadrant (rows of cta0/cta1 x B columns of cta0/cta1): [%g %g] [%g %g]\n", worst[0][0], worst[0][1], worst[1][0], worst[1][1]);
  const bool ok = worst[0][0] == 0 && worst[0][1] == 0 && worst[1][0] == 0 && worst[1][1] == 0;
  printf("%s\n", ok ? "PAIR PROBE OK" : "PAIR PROBE MISMATCH");
  if (!ok) { printf("D[0][0..3] = %g %g %g %g ; D[128][0..3] = %g %g %g %g ; D[0][128..131] = %g %g %g %g\n", Dh[0], Dh[1], Dh[2], Dh[3],
                    Dh[128 * 256], Dh[128 * 256 + 1], Dh[128 * 256 + 2], Dh[128 * 256 + 3], Dh[128], Dh[129], Dh[130], Dh[131]); }
  return ok ? 0 : 1;
}
