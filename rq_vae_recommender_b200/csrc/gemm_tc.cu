// bf16 tcgen05 GEMM with fused ReLU for the encoder / decoder MLPs (modules/encoder.py:23-38), sm_100a.
//
//   Y[M,N] = act( X[M,K] . W[N,K]^T ),   bf16 operands, fp32 accumulation in TMEM, act = ReLU or identity.
//
// This is the reduced-precision (AMP-like) path: the reference runs these Linears in bf16 when
// `train_rqvae.py:36,69` enables mixed precision.  The exact fp32 path (csrc/dense.cu sgemm) stays the default because
// index parity at 1e-5 needs it; this kernel is opt-in and forward-only (tokenisation).
//
// Data layout ("image"): every operand is stored in HBM as the exact shared-memory image the tensor core reads --
// [row-tile of 128][k-chunk of 64][128 rows x 128 B], K-major, 16-byte chunks XOR-swizzled with (row & 7) (UMMA
// SWIZZLE_128B).  A stage is then ONE contiguous 16 KB TMA bulk copy, no tensor maps, and the epilogue of layer i
// writes layer i+1's A operand directly in that layout, so activations never exist in row-major form.
//
// Kernel: persistent, one CTA per SM.  warp 0 = TMA producer (A + up to 2 W blocks per stage, 4-stage ring),
// warp 1 = MMA issuer (tcgen05.mma M128 N128 K16, accumulators double-buffered in TMEM: 2 x 256 columns),
// warps 4-11 = epilogue (2 per TMEM lane quarter, one per 128-column block): tcgen05.ld -> ReLU -> bf16 image or fp32 rows.
#include "common.cuh"
#include <cuda_bf16.h>

#define GT_KC 64
#define GT_BLK_BYTES (128 * GT_KC * 2)   // 16 KB: 128 rows x 64 bf16
#define GT_STAGES 4
#define GT_THREADS 384                   // warps 0-3: producer, MMA, 2 idle | warps 4-11: epilogue

// ------------------------------------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void gt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gt_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void gt_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void gt_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void gt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gt_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B smem descriptor / instruction descriptor: see rq_tc.cu (same encodings); A,B format 1 = BF16
__device__ __forceinline__ uint64_t gt_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t gt_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ image builders
extern "C" size_t rqb200_bf16_image_bytes(int rows, int K) {
  if (rows < 0 || K <= 0 || K % GT_KC) return 0;
  return (size_t)((rows + 127) / 128) * (K / GT_KC) * GT_BLK_BYTES;
}

// fp32 row-major [rows, K] -> bf16 image; rows beyond `rows` in the last tile are zero.  One CTA per (row tile, k chunk).
__global__ void gt_f32_to_image_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, __nv_bfloat16* img) {
  const int nkc = K / GT_KC;
  const int mt = blockIdx.x / nkc, kc = blockIdx.x % nkc;
  unsigned char* out = reinterpret_cast<unsigned char*>(img) + (size_t)blockIdx.x * GT_BLK_BYTES;
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;              // row in tile, 16-byte chunk (8 elements)
    const int row = mt * 128 + r;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (row < rows) {
      const float* src = x + (int64_t)row * ldx + kc * GT_KC + c * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
    }
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 w;
    w.x = *reinterpret_cast<uint32_t*>(&p0); w.y = *reinterpret_cast<uint32_t*>(&p1);
    w.z = *reinterpret_cast<uint32_t*>(&p2); w.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(out + r * 128 + ((c ^ (r & 7)) << 4)) = w;
  }
}

extern "C" int rqb200_f32_to_bf16_image(const float* x, int64_t ldx, int rows, int K, void* image, void* stream) {
  RQB_CHECK_ARG(K > 0 && K % GT_KC == 0 && rows >= 0 && ldx >= K, "f32_to_bf16_image: need K %% 64 == 0 (K=%d)", K);
  if (rows == 0) return RQB_OK;
  RQB_CHECK_ARG(x && image, "f32_to_bf16_image: null pointer");
  const int blocks = ((rows + 127) / 128) * (K / GT_KC);
  gt_f32_to_image_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, ldx, rows, K, reinterpret_cast<__nv_bfloat16*>(image));
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM
struct GtParams {
  const unsigned char* a_img;   // [mtiles][nkc][16 KB]
  const unsigned char* w_img;   // [nblocks][nkc][16 KB]   (rows of W padded with zeros to a multiple of 128)
  int M, N, K, nkc, mtiles, nblocks, ngroups, nitems;
  int relu;
  unsigned char* out_img;       // next layer's A image [mtiles][N/64][16 KB] (N % 64 == 0), or null
  float* out_f32;               // row-major [M, N] (ld = ldo), or null
  int64_t ldo;
};

struct GtSmemMisc {
  uint64_t full[GT_STAGES], empty[GT_STAGES];
  uint64_t t_full[2][2], t_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__global__ void __launch_bounds__(GT_THREADS, 1) gt_gemm_kernel(GtParams p) {
  extern __shared__ __align__(1024) unsigned char gsm[];
  // stage s: [A 16 KB][W block 0 16 KB][W block 1 16 KB]
  GtSmemMisc* ms = reinterpret_cast<GtSmemMisc*>(gsm + GT_STAGES * 3 * GT_BLK_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(gsm) & 1023u) != 0) __trap();
    for (int i = 0; i < GT_STAGES; ++i) { mbar_init(&ms->full[i], 1); mbar_init(&ms->empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ms->t_full[i][0], 1); mbar_init(&ms->t_full[i][1], 1);
      mbar_init(&ms->t_empty[i], 8 * 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) gt_alloc(&ms->tmem_base, 512);
  gt_fence_before();
  __syncthreads();
  gt_fence_after();
#define GT_TMEM() (*reinterpret_cast<volatile uint32_t*>(&ms->tmem_base))

  // work item = (row tile mt, group of up to two 128-column W blocks)
  if (warp == 0 && lane == 0) {
    // ============================================================== TMA producer
    uint32_t s = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int mt = item / p.ngroups, g = item % p.ngroups;
      const int nb = min(2, p.nblocks - 2 * g);
      for (int kc = 0; kc < p.nkc; ++kc, ++s) {
        const uint32_t st = s % GT_STAGES, u = s / GT_STAGES;
        mbar_wait_guarded(&ms->empty[st], (u & 1) ^ 1, 1);
        unsigned char* dst = gsm + st * 3 * GT_BLK_BYTES;
        mbar_expect_tx(&ms->full[st], (1 + nb) * GT_BLK_BYTES);
        bulk_g2s(dst, p.a_img + ((size_t)mt * p.nkc + kc) * GT_BLK_BYTES, GT_BLK_BYTES, &ms->full[st]);
        for (int b = 0; b < nb; ++b)
          bulk_g2s(dst + (1 + b) * GT_BLK_BYTES, p.w_img + ((size_t)(2 * g + b) * p.nkc + kc) * GT_BLK_BYTES, GT_BLK_BYTES,
                   &ms->full[st]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ============================================================== MMA issuer
    const uint32_t idesc = gt_idesc_bf16(128, 128);
    const uint32_t base = smem_u32(gsm);
    uint32_t s = 0, it = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      const int g = item % p.ngroups;
      const int nb = min(2, p.nblocks - 2 * g);
      const uint32_t buf = it & 1, u = it >> 1;
      mbar_wait_guarded(&ms->t_empty[buf], (u & 1) ^ 1, 2);
      gt_fence_after();
      for (int kc = 0; kc < p.nkc; ++kc, ++s) {
        const uint32_t st = s % GT_STAGES;
        mbar_wait_guarded(&ms->full[st], (s / GT_STAGES) & 1, 3);
        gt_fence_after();
        const uint32_t sa = base + st * 3 * GT_BLK_BYTES;
        const uint64_t adesc = gt_smem_desc(sa);
        for (int b = 0; b < nb; ++b) {
          const uint64_t bdesc = gt_smem_desc(sa + (1 + b) * GT_BLK_BYTES);
          const uint32_t d = GT_TMEM() + buf * 256 + b * 128;
#pragma unroll
          for (int j = 0; j < GT_KC / 16; ++j) gt_mma(d, adesc + 2 * j, bdesc + 2 * j, idesc, (kc | j) != 0);
        }
        gt_commit(&ms->empty[st]);
      }
      gt_commit(&ms->t_full[buf][0]);
      gt_commit(&ms->t_full[buf][1]);
    }
  } else if (warp >= 4) {
    // ============================================================== epilogue: TMEM -> act -> bf16 image / fp32 rows
    const int quarter = warp & 3, blk = (warp - 4) >> 2;      // TMEM lane quarter, 128-column block inside the group
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      const int mt = item / p.ngroups, g = item % p.ngroups;
      const int nb = min(2, p.nblocks - 2 * g);
      const uint32_t buf = it & 1, u = it >> 1;
      mbar_wait_guarded(&ms->t_full[buf][blk], u & 1, 4);
      gt_fence_after();
      const int row = mt * 128 + r;
      const int col0 = (2 * g + blk) * 128;                    // first output column of this block
      if (blk < nb) {
        const uint32_t tcol = GT_TMEM() + lane_addr + buf * 256 + blk * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = col0 + c * 32;
          if (col >= p.N) break;                               // warp-uniform
          uint32_t sr[32];
          gt_ld32(tcol + c * 32, sr);
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            v[e] = __uint_as_float(sr[e]);
            if (p.relu) v[e] = fmaxf(v[e], 0.f);
          }
          if (p.out_img) {
            // columns [col, col+32) = half of k-chunk (col / 64) of the next layer's A image: 4 x 16-byte swizzled chunks
            unsigned char* dst = p.out_img + ((size_t)mt * (p.N / GT_KC) + (col >> 6)) * GT_BLK_BYTES + r * 128;
            const int cbase = (col & 63) >> 3;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]), p1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]), p3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
              uint4 w;
              w.x = *reinterpret_cast<uint32_t*>(&p0); w.y = *reinterpret_cast<uint32_t*>(&p1);
              w.z = *reinterpret_cast<uint32_t*>(&p2); w.w = *reinterpret_cast<uint32_t*>(&p3);
              *reinterpret_cast<uint4*>(dst + (((cbase + q) ^ (r & 7)) << 4)) = w;   // rows >= M hold act(0) = 0: harmless padding
            }
          }
          if (p.out_f32 && row < p.M) {
            float* o = p.out_f32 + (int64_t)row * p.ldo + col;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col + e < p.N) o[e] = v[e];
          }
        }
      }
      gt_fence_before();
      mbar_arrive(&ms->t_empty[buf]);
    }
  }

  gt_fence_before();
  __syncthreads();
  if (warp == 1) {
    gt_fence_after();
    gt_dealloc(GT_TMEM(), 512);
  }
}

extern "C" int rqb200_gemm_bf16(const void* a_image, const void* w_image, int M, int N, int K, int relu, void* out_image,
                                float* out_f32, int64_t ldo, void* stream) {
  RQB_CHECK_ARG(M >= 0 && N > 0 && K > 0 && K % GT_KC == 0, "gemm_bf16: need K %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  RQB_CHECK_ARG(!out_image || N % GT_KC == 0, "gemm_bf16: an image output needs N %% 64 == 0 (N=%d)", N);
  RQB_CHECK_ARG(out_image || out_f32, "gemm_bf16: no output");
  RQB_CHECK_ARG(!out_f32 || ldo >= N, "gemm_bf16: ldo < N");
  if (M == 0) return RQB_OK;
  RQB_CHECK_ARG(a_image && w_image, "gemm_bf16: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GtParams p{};
  p.a_img = reinterpret_cast<const unsigned char*>(a_image);
  p.w_img = reinterpret_cast<const unsigned char*>(w_image);
  p.M = M; p.N = N; p.K = K; p.nkc = K / GT_KC;
  p.mtiles = (M + 127) / 128;
  p.nblocks = (N + 127) / 128;
  p.ngroups = (p.nblocks + 1) / 2;
  p.nitems = p.mtiles * p.ngroups;
  p.relu = relu;
  p.out_img = reinterpret_cast<unsigned char*>(out_image);
  p.out_f32 = out_f32; p.ldo = ldo;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    RQB_CUDA(cudaGetDevice(&dev));
    RQB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t smem = (size_t)GT_STAGES * 3 * GT_BLK_BYTES + sizeof(GtSmemMisc);
  RQB_CUDA(cudaFuncSetAttribute(gt_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.nitems < sm_count ? p.nitems : sm_count;
  gt_gemm_kernel<<<grid, GT_THREADS, smem, st>>>(p);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}
