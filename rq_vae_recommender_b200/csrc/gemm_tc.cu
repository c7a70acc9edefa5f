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
  int dev = 0, sm_count = 0;                                 // per call: the current device may differ between calls
  RQB_CUDA(cudaGetDevice(&dev));
  RQB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = (size_t)GT_STAGES * 3 * GT_BLK_BYTES + sizeof(GtSmemMisc);
  RQB_CUDA(cudaFuncSetAttribute(gt_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.nitems < sm_count ? p.nitems : sm_count;
  gt_gemm_kernel<<<grid, GT_THREADS, smem, st>>>(p);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

// =====================================================================================================================
// Split-precision GEMM: fp32-accurate products on the fp16 tensor cores (the MLPs of modules/encoder.py:23-38 in their
// default, index-exact precision; the two GEMMs of a Gumbel-softmax level, modules/quantize.py:113-117,135).
//
//   C[M,N] = act( A[M,K] . B[N,K]^T ),  A, B, C fp32 in HBM.
//
// Every operand row is scaled by a power of two so that its largest element lies in [2^14, 2^15) and stored as TWO fp16 images
//   hi = fp16(v 2^e),  lo = fp16(v 2^e - hi)          (hi + lo carries 22 significant bits of v; lo may be subnormal: the error
//                                                      is then 2^-25 absolute = 2^-39 of the row maximum)
// and the product is three tcgen05.mma per k-step with fp32 accumulation in TMEM:  hi.hi + lo.hi + hi.lo  (lo.lo is 2^-22
// relative and dropped).  The epilogue multiplies by 2^-(e_row + e_col), both exact.  Measured against float64 the result is
// as close as a plain fp32 FMA GEMM (tests/test_gpu_gemm_split.py states the bound that is asserted).
//
// Image = the bf16 image's layout with fp16 elements: [row tile of 128][k chunk of 64][128 rows x 128 B swizzled]; one buffer
// holds [hi image][lo image][row scales: 128 floats per row tile, value 2^-e].  K is padded with zeros to a multiple of 64, rows
// to a multiple of 128.
//
// Kernel: persistent, one CTA per SM; work item = (row tile, group of up to 256 columns).  Stage = [A hi][A lo][B hi 0][B hi 1]
// [B lo 0][B lo 1] = 96 KB, 2 stages; the two B blocks of a half are adjacent so that ONE tcgen05.mma covers N = 256.
// warp 0 = bulk-copy producer, warp 1 = MMA issuer (two accumulators of 256 TMEM columns: hi.hi and the cross terms),
// warps 4-11 = epilogue (sums the two, scales, activates).
#include <cuda_fp16.h>

#define GS_STAGES 2
#define GS_MAX_CHUNKS 12                   // 16-byte chunks per lane of the row splitter: K <= 32 * 8 * 12 = 3072
#define GS_STAGE_BYTES (6 * GT_BLK_BYTES)

extern "C" size_t rqb200_split_image_bytes(int rows, int K) {
  if (rows < 0 || K <= 0) return 0;
  const size_t mt = (size_t)(rows + 127) / 128, nkc = (size_t)(K + GT_KC - 1) / GT_KC;
  return 2 * mt * nkc * GT_BLK_BYTES + mt * 128 * sizeof(float);
}

__device__ __forceinline__ float gs_pow2_scale(float mx) {
  // 2^e with mx 2^e in [2^14, 2^15); zero / non-finite rows are not scaled
  if (!(mx > 0.f) || !(mx < INFINITY)) return 1.f;
  int ex;
  frexpf(mx, &ex);                        // mx = f 2^ex, f in [0.5, 1)
  return ldexpf(1.f, max(-120, min(120, 15 - ex)));      // (clamped: 2^e and 2^-e both stay normal)
}
__device__ __forceinline__ void gs_split8(const float (&v)[8], float s, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = v[2 * e] * s, b = v[2 * e + 1] * s;
    h[e] = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h[e]);
    l[e] = __floats2half2_rn(a - hf.x, b - hf.y);
  }
  hi.x = *reinterpret_cast<uint32_t*>(&h[0]); hi.y = *reinterpret_cast<uint32_t*>(&h[1]);
  hi.z = *reinterpret_cast<uint32_t*>(&h[2]); hi.w = *reinterpret_cast<uint32_t*>(&h[3]);
  lo.x = *reinterpret_cast<uint32_t*>(&l[0]); lo.y = *reinterpret_cast<uint32_t*>(&l[1]);
  lo.z = *reinterpret_cast<uint32_t*>(&l[2]); lo.w = *reinterpret_cast<uint32_t*>(&l[3]);
}

// Row-major source [rows, K] (ld = ldx): W lanes per image row (32 / W rows per warp pass), NCH 16-byte chunks per lane; the row
// stays in registers between the maximum and the split.  grid = row tiles, block = 256.  Few registers on purpose: the kernel
// is a pure HBM stream (read 4 B, write 4 B per element) and needs many warps per SM in flight -- the first version (one
// generic 12-chunk instantiation, 133 registers, one CTA per SM) ran at 0.8-2 TB/s.
template <int NCH, int W>
__global__ void __launch_bounds__(256) gs_split_rows_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, unsigned char* img) {
  const int nkc = (K + GT_KC - 1) / GT_KC, mtiles = gridDim.x;
  const int mt = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* hi_img = img;
  unsigned char* lo_img = img + (size_t)mtiles * nkc * GT_BLK_BYTES;
  float* scales = reinterpret_cast<float*>(img + 2 * (size_t)mtiles * nkc * GT_BLK_BYTES);
  const int nchunks = nkc * 8;                                    // 16-byte (8 element) chunks per image row
  const bool vec = (K % 8 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  constexpr int G = 32 / W;                                       // rows per warp pass
  const int sub = lane / W, sl = lane % W;
#pragma unroll 1
  for (int r = warp * G + sub; r < 128; r += 8 * G) {
    const int row = mt * 128 + r;
    float v[NCH][8];
    float mx = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + W * i;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
      if (c < nchunks && row < rows) {
        const float* src = x + (int64_t)row * ldx + c * 8;
        if (vec && c * 8 + 8 <= K) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (c * 8 + e < K) v[i][e] = __ldg(src + e);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, fabsf(v[i][e]));
      }
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float s = gs_pow2_scale(mx);
    if (sl == 0) scales[mt * 128 + r] = 1.f / s;                   // exact: a power of two
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = sl + W * i;
      if (c < nchunks) {
        uint4 hi, lo;
        gs_split8(v[i], s, hi, lo);
        const size_t off = ((size_t)mt * nkc + (c >> 3)) * GT_BLK_BYTES + r * 128 + (((c & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(hi_img + off) = hi;
        *reinterpret_cast<uint4*>(lo_img + off) = lo;
      }
    }
  }
}

// Transposed source: image row r = column r of x[K, rows] (ld = ldx) -- the operand of x^T without materialising the transpose
// (W^T for dgrad, C^T for W@C, g^T and h^T for the weight gradients whose contraction runs over the batch).  Three launches:
//   gs_colmax_kernel    |column| maxima by atomicMax on the float bits (non-negative floats order like unsigned ints) into scales[]
//   gs_colscale_kernel  scales[r] = 2^-e
//   gs_split_cols_kernel one CTA per (row tile, k chunk): 64 x 128 source floats through shared memory (coalesced reads, the
//                       transposed reads are conflict-free with a 129-float pitch), hi / lo blocks written as 16-byte chunks
__global__ void gs_colmax_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, unsigned int* colmax) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int k0 = blockIdx.y * 128;
  if (c >= rows) return;
  float mx = 0.f;
  const int k1 = min(K, k0 + 128);
  for (int k = k0; k < k1; ++k) mx = fmaxf(mx, fabsf(__ldg(x + (int64_t)k * ldx + c)));
  if (mx != mx) mx = INFINITY;                               // NaN: not scaled (gs_pow2_scale), like the row kernel
  atomicMax(colmax + c, __float_as_uint(mx));
}
__global__ void gs_colscale_kernel(float* scales, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scales[i] = 1.f / gs_pow2_scale(scales[i]);
}
__global__ void __launch_bounds__(256) gs_split_cols_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, unsigned char* img) {
  __shared__ float tile[GT_KC][129];
  const int mtiles = gridDim.x, nkc = gridDim.y;
  const int mt = blockIdx.x, kc = blockIdx.y, t = threadIdx.x;
  unsigned char* hi_img = img;
  unsigned char* lo_img = img + (size_t)mtiles * nkc * GT_BLK_BYTES;
  const float* scales = reinterpret_cast<const float*>(img + 2 * (size_t)mtiles * nkc * GT_BLK_BYTES);
#pragma unroll 4
  for (int i = 0; i < (GT_KC * 128) / 256; ++i) {
    const int idx = t + 256 * i, kr = idx >> 7, c = idx & 127;
    const int k = kc * GT_KC + kr, row = mt * 128 + c;
    tile[kr][c] = (k < K && row < rows) ? __ldg(x + (int64_t)k * ldx + row) : 0.f;
  }
  __syncthreads();
  const int r = t & 127;
  const float s = 1.f / scales[mt * 128 + r];                 // exact: a power of two
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int c = (t >> 7) * 4 + cc;                          // 16-byte chunk of the 128-byte block row
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[c * 8 + e][r];
    uint4 hi, lo;
    gs_split8(v, s, hi, lo);
    const size_t off = ((size_t)mt * nkc + kc) * GT_BLK_BYTES + r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(hi_img + off) = hi;
    *reinterpret_cast<uint4*>(lo_img + off) = lo;
  }
}

extern "C" int rqb200_f32_to_split_image(const float* x, int64_t ldx, int rows, int K, int transposed, void* image, void* stream) {
  RQB_CHECK_ARG(K > 0 && rows >= 0, "f32_to_split_image: bad shape (rows=%d K=%d)", rows, K);
  RQB_CHECK_ARG(transposed ? ldx >= rows : ldx >= K, "f32_to_split_image: ld too small");
  if (rows == 0) return RQB_OK;
  RQB_CHECK_ARG(x && image, "f32_to_split_image: null pointer");
  RQB_CHECK_ARG((reinterpret_cast<uintptr_t>(image) & 15) == 0, "f32_to_split_image: the image must be 16-byte aligned");
  const int mtiles = (rows + 127) / 128;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (transposed) {
    unsigned char* im = reinterpret_cast<unsigned char*>(image);
    const int nkc = (K + GT_KC - 1) / GT_KC;
    if (nkc > 65535) {                                       // grid.y of the column kernels
      rqb_set_error("f32_to_split_image: transposed operand with K = %d > %d", K, 65535 * GT_KC);
      return RQB_ERR_UNSUPPORTED;
    }
    float* scales = reinterpret_cast<float*>(im + 2 * (size_t)mtiles * nkc * GT_BLK_BYTES);
    RQB_CUDA(cudaMemsetAsync(scales, 0, (size_t)mtiles * 128 * sizeof(float), st));
    gs_colmax_kernel<<<dim3((rows + 255) / 256, (K + 127) / 128), 256, 0, st>>>(x, ldx, rows, K, reinterpret_cast<unsigned int*>(scales));
    RQB_LAUNCH_CHECK();
    gs_colscale_kernel<<<(mtiles * 128 + 255) / 256, 256, 0, st>>>(scales, mtiles * 128);
    RQB_LAUNCH_CHECK();
    gs_split_cols_kernel<<<dim3(mtiles, nkc), 256, 0, st>>>(x, ldx, rows, K, im);
  } else {
    if (K > 32 * 8 * GS_MAX_CHUNKS) {
      rqb_set_error("f32_to_split_image: K = %d > %d", K, 32 * 8 * GS_MAX_CHUNKS);
      return RQB_ERR_UNSUPPORTED;
    }
    unsigned char* im = reinterpret_cast<unsigned char*>(image);
    const int nchunks = ((K + GT_KC - 1) / GT_KC) * 8;
    if (nchunks <= 8) gs_split_rows_kernel<1, 8><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
    else if (nchunks <= 16) gs_split_rows_kernel<1, 16><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
    else if (nchunks <= 32) gs_split_rows_kernel<1, 32><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
    else if (nchunks <= 64) gs_split_rows_kernel<2, 32><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
    else if (nchunks <= 96) gs_split_rows_kernel<3, 32><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
    else if (nchunks <= 128) gs_split_rows_kernel<4, 32><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
    else gs_split_rows_kernel<GS_MAX_CHUNKS, 32><<<mtiles, 256, 0, st>>>(x, ldx, rows, K, im);
  }
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

struct GsParams {
  const unsigned char *a_hi, *a_lo, *b_hi, *b_lo;   // [tiles][nkc][16 KB]
  const float *a_scale, *b_scale;                   // 2^-e per image row
  int M, N, nkc, mtiles, nblocks, ngroups, nitems, relu;
  int ksplit, kc_per;                               // split-K: item = (row tile, column group, k slice of kc_per chunks); slice ks
  int64_t part_stride;                              // writes its partial sums to out + ks * part_stride (ksplit == 1: 0)
  float* out;
  int64_t ldo;
  const float* mask;                                // optional [M, N] (ld = ldm): out = mask > 0 ? out : 0  (ReLU' of a backward GEMM)
  int64_t ldm;
};

__device__ __forceinline__ uint32_t gs_idesc_f16(int M, int N) {       // A, B = F16 (format 0), D = F32, K-major both
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(GT_THREADS, 1) gs_gemm_kernel(GsParams p) {
  extern __shared__ __align__(1024) unsigned char gsm[];
  GtSmemMisc* ms = reinterpret_cast<GtSmemMisc*>(gsm + GS_STAGES * GS_STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    if ((smem_u32(gsm) & 1023u) != 0) __trap();
    for (int i = 0; i < GS_STAGES; ++i) { mbar_init(&ms->full[i], 1); mbar_init(&ms->empty[i], 1); }
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

  if (warp == 0 && lane == 0) {
    // ============================================================== producer: 2 A blocks + 2 or 4 B blocks per stage
    uint32_t s = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
      const int ks = item % p.ksplit, tile = item / p.ksplit;
      const int mt = tile / p.ngroups, g = tile % p.ngroups;
      const int nb = min(2, p.nblocks - 2 * g);
      const int kc_end = min(p.nkc, (ks + 1) * p.kc_per);
      for (int kc = ks * p.kc_per; kc < kc_end; ++kc, ++s) {
        const uint32_t st = s % GS_STAGES, u = s / GS_STAGES;
        mbar_wait_guarded(&ms->empty[st], (u & 1) ^ 1, 1);
        unsigned char* dst = gsm + st * GS_STAGE_BYTES;
        mbar_expect_tx(&ms->full[st], (2 + 2 * nb) * GT_BLK_BYTES);
        const size_t ao = ((size_t)mt * p.nkc + kc) * GT_BLK_BYTES;
        bulk_g2s(dst, p.a_hi + ao, GT_BLK_BYTES, &ms->full[st]);
        bulk_g2s(dst + GT_BLK_BYTES, p.a_lo + ao, GT_BLK_BYTES, &ms->full[st]);
        for (int b = 0; b < nb; ++b) {
          const size_t bo = ((size_t)(2 * g + b) * p.nkc + kc) * GT_BLK_BYTES;
          bulk_g2s(dst + (2 + b) * GT_BLK_BYTES, p.b_hi + bo, GT_BLK_BYTES, &ms->full[st]);
          bulk_g2s(dst + (4 + b) * GT_BLK_BYTES, p.b_lo + bo, GT_BLK_BYTES, &ms->full[st]);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ============================================================== MMA issuer: hi.hi + lo.hi + hi.lo per k-step
    const uint32_t base = smem_u32(gsm);
    uint32_t s = 0, it = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      const int ks = item % p.ksplit, g = (item / p.ksplit) % p.ngroups;
      const int nb = min(2, p.nblocks - 2 * g);
      const uint32_t idesc = gs_idesc_f16(128, nb * 128);
      const int kc_begin = ks * p.kc_per, kc_end = min(p.nkc, (ks + 1) * p.kc_per);
      mbar_wait_guarded(&ms->t_empty[0], (it & 1) ^ 1, 2);
      gt_fence_after();
      // two accumulators: hi.hi in columns [0, 256), the cross terms in [256, 512).  The tensor core truncates the fp32
      // accumulator after every MMA (measured: the error of one shared accumulator grows with the number of MMAs and matches a
      // round-toward-zero model, 5.4e-7 of |a||b| at K = 768); apart, the large sum sees a third of the truncations and the
      // small one truncates at 2^-11 of the magnitude: 1.8e-7, the level of a plain fp32 GEMM
      const uint32_t d = GT_TMEM(), dx = GT_TMEM() + 256;
      for (int kc = kc_begin; kc < kc_end; ++kc, ++s) {
        const uint32_t st = s % GS_STAGES;
        mbar_wait_guarded(&ms->full[st], (s / GS_STAGES) & 1, 3);
        gt_fence_after();
        const uint32_t sa = base + st * GS_STAGE_BYTES;
        const uint64_t ahi = gt_smem_desc(sa), alo = gt_smem_desc(sa + GT_BLK_BYTES);
        const uint64_t bhi = gt_smem_desc(sa + 2 * GT_BLK_BYTES), blo = gt_smem_desc(sa + 4 * GT_BLK_BYTES);
#pragma unroll
        for (int j = 0; j < GT_KC / 16; ++j) {
          gt_mma(d, ahi + 2 * j, bhi + 2 * j, idesc, ((kc - kc_begin) | j) != 0);
          gt_mma(dx, alo + 2 * j, bhi + 2 * j, idesc, ((kc - kc_begin) | j) != 0);
          gt_mma(dx, ahi + 2 * j, blo + 2 * j, idesc, 1);
        }
        gt_commit(&ms->empty[st]);
      }
      gt_commit(&ms->t_full[0][0]);
      gt_commit(&ms->t_full[0][1]);
    }
  } else if (warp >= 4) {
    // ============================================================== epilogue: TMEM -> x 2^-(e_row + e_col) -> act -> fp32 rows
    const int quarter = warp & 3, blk = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it) {
      const int ks = item % p.ksplit, tile = item / p.ksplit;
      const int mt = tile / p.ngroups, g = tile % p.ngroups;
      const int nb = min(2, p.nblocks - 2 * g);
      mbar_wait_guarded(&ms->t_full[0][blk], it & 1, 4);
      gt_fence_after();
      const int row = mt * 128 + r;
      const int col0 = (2 * g + blk) * 128;
      if (blk < nb) {
        const float rs = __ldg(p.a_scale + row);                 // scale vectors are padded to whole tiles
        const uint32_t tcol = GT_TMEM() + lane_addr + blk * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int col = col0 + c * 32;
          if (col >= p.N) break;                                 // warp-uniform
          uint32_t sr[32], sx[32];
          gt_ld32(tcol + c * 32, sr);
          gt_ld32(tcol + 256 + c * 32, sx);
          // scale of column col + lane (vectors are padded to whole tiles).  2^-(e_row + e_col) is applied as two factors of half
          // the exponent each (|e| <= 120 per operand): neither the factor nor the intermediate product leaves the fp32 range
          // unless the result does
          const int ea = (__float_as_int(rs) >> 23) & 0xff;                                    // this lane's ROW
          const int ebl = (__float_as_int(__ldg(p.b_scale + col + lane)) >> 23) & 0xff;         // column col + lane
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int et = ea + __shfl_sync(0xffffffffu, ebl, e) - 254;                         // -(e_row + e_col)
            const int e1 = et >> 1;
            v[e] = ((__uint_as_float(sr[e]) + __uint_as_float(sx[e])) * __int_as_float((e1 + 127) << 23)) * __int_as_float((et - e1 + 127) << 23);   // exact
            if (p.relu) v[e] = fmaxf(v[e], 0.f);
          }
          if (p.mask && row < p.M) {
            const float* mk = p.mask + (int64_t)row * p.ldm + col;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col + e < p.N && !(__ldg(mk + e) > 0.f)) v[e] = 0.f;
          }
          if (row < p.M) {
            float* o = p.out + (int64_t)ks * p.part_stride + (int64_t)row * p.ldo + col;
            if (col + 32 <= p.N && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
#pragma unroll
              for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(o + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (col + e < p.N) o[e] = v[e];
            }
          }
        }
      }
      gt_fence_before();
      mbar_arrive(&ms->t_empty[0]);
    }
  }

  gt_fence_before();
  __syncthreads();
  if (warp == 1) {
    gt_fence_after();
    gt_dealloc(GT_TMEM(), 512);
  }
}

// out[i, j] = sum over the k slices of the partial sums (fixed order: deterministic)
__global__ void gs_reduce_kernel(const float* __restrict__ part, int S, int M, int N, float* __restrict__ out, int64_t ldo) {
  const int64_t n = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += part[(int64_t)s * n + i];
    out[(i / N) * ldo + (i % N)] = acc;
  }
}

static int gs_run(const void* a_image, const void* b_image, int M, int N, int K, int relu, const float* mask, int64_t ldm,
                  float* out, int64_t ldo, int ksplit, int kc_per, int64_t part_stride, cudaStream_t st) {
  GsParams p{};
  p.M = M; p.N = N; p.nkc = (K + GT_KC - 1) / GT_KC;
  p.mtiles = (M + 127) / 128;
  p.nblocks = (N + 127) / 128;
  p.ngroups = (p.nblocks + 1) / 2;
  p.ksplit = ksplit; p.kc_per = kc_per; p.part_stride = part_stride;
  p.nitems = p.mtiles * p.ngroups * ksplit;
  p.relu = relu;
  const size_t a_img = (size_t)p.mtiles * p.nkc * GT_BLK_BYTES, b_img = (size_t)p.nblocks * p.nkc * GT_BLK_BYTES;
  p.a_hi = reinterpret_cast<const unsigned char*>(a_image); p.a_lo = p.a_hi + a_img;
  p.a_scale = reinterpret_cast<const float*>(p.a_hi + 2 * a_img);
  p.b_hi = reinterpret_cast<const unsigned char*>(b_image); p.b_lo = p.b_hi + b_img;
  p.b_scale = reinterpret_cast<const float*>(p.b_hi + 2 * b_img);
  p.out = out; p.ldo = ldo;
  p.mask = mask; p.ldm = ldm;
  int dev = 0, sm_count = 0;
  RQB_CUDA(cudaGetDevice(&dev));
  RQB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = (size_t)GS_STAGES * GS_STAGE_BYTES + sizeof(GtSmemMisc);
  RQB_CUDA(cudaFuncSetAttribute(gs_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.nitems < sm_count ? p.nitems : sm_count;
  gs_gemm_kernel<<<grid, GT_THREADS, smem, st>>>(p);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

extern "C" int rqb200_gemm_split(const void* a_image, const void* b_image, int M, int N, int K, int relu, const float* mask,
                                 int64_t ldm, float* out, int64_t ldo, void* stream) {
  RQB_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldo >= N, "gemm_split: bad shape (M=%d N=%d K=%d ldo=%lld)", M, N, K, (long long)ldo);
  if (M == 0) return RQB_OK;
  RQB_CHECK_ARG(a_image && b_image && out, "gemm_split: null pointer");
  RQB_CHECK_ARG(!mask || ldm >= N, "gemm_split: ldm < N");
  return gs_run(a_image, b_image, M, N, K, relu, mask, ldm, out, ldo, 1, (K + GT_KC - 1) / GT_KC, 0,
                reinterpret_cast<cudaStream_t>(stream));
}

// Split-K schedule for products with few output tiles and a long contraction (the weight gradients: M = out, N = in, K = batch):
// the k chunks are cut into `slices` ranges, every (tile, range) is a work item writing partial sums into the workspace
// [slices][M][N], and a fixed-order reduction produces out.  gemm_split_k_slices picks the slice count that fills the SMs.
extern "C" int rqb200_gemm_split_k_slices(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 1;
  int dev = 0, sm_count = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  const int nkc = (K + GT_KC - 1) / GT_KC;
  const int tiles = ((M + 127) / 128) * (((N + 127) / 128 + 1) / 2);
  int want = (sm_count + tiles - 1) / tiles;                 // slices that give every SM an item
  if (want > nkc / 4) want = nkc / 4;                        // at least 4 chunks (256 k) per slice
  if (want < 1) want = 1;
  const int kc_per = (nkc + want - 1) / want;
  return (nkc + kc_per - 1) / kc_per;                        // no empty slice
}

extern "C" int rqb200_gemm_split_k(const void* a_image, const void* b_image, int M, int N, int K, int slices, float* workspace,
                                   float* out, int64_t ldo, void* stream) {
  RQB_CHECK_ARG(M >= 0 && N > 0 && K > 0 && ldo >= N && slices >= 1, "gemm_split_k: bad shape (M=%d N=%d K=%d slices=%d)", M, N, K, slices);
  if (M == 0) return RQB_OK;
  RQB_CHECK_ARG(a_image && b_image && out, "gemm_split_k: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nkc = (K + GT_KC - 1) / GT_KC;
  const int kc_per = (nkc + slices - 1) / slices;
  RQB_CHECK_ARG((int64_t)(slices - 1) * kc_per < nkc, "gemm_split_k: %d slices of %d chunks leave an empty slice (K = %d)", slices, kc_per, K);
  if (slices == 1) return gs_run(a_image, b_image, M, N, K, 0, nullptr, 0, out, ldo, 1, nkc, 0, st);
  RQB_CHECK_ARG(workspace, "gemm_split_k: null workspace");
  int rc = gs_run(a_image, b_image, M, N, K, 0, nullptr, 0, workspace, N, slices, kc_per, (int64_t)M * N, st);
  if (rc) return rc;
  const int64_t n = (int64_t)M * N;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  gs_reduce_kernel<<<grid, 256, 0, st>>>(workspace, slices, M, N, out, ldo);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}
