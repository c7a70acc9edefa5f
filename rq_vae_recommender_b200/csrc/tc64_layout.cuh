// Index arithmetic of rq_tc64_kernel (csrc/rq_tc64.cu), written so the SAME code compiles for the device (nvcc) and for the
// host (g++: tests/test_tc64_layout.py checks it against the layouts' definitions).  Keeping these few expressions out of
// the kernel body is what lets them be tested without a GPU.
#ifndef RQB200_TC64_LAYOUT_CUH
#define RQB200_TC64_LAYOUT_CUH
#include <stdint.h>

#ifdef __CUDACC__
#define TC64_HD __host__ __device__ __forceinline__
#else
#define TC64_HD inline
#endif

// fp32 staging stage written by TMA (box = 64 rows x 64 floats of x, no swizzle): byte offset of float4 column q (0..15) of
// row r (0..63).  Row-major, 256 bytes per row.
TC64_HD uint32_t tc64_stage_offset(int r, int q) { return (uint32_t)r * 256u + (uint32_t)q * 16u; }

// A chunk (64 rows x 64 fp16, K-major SWIZZLE_128B as tcgen05.mma reads it: 128 bytes per row, the eight 16-byte units of a
// row XOR-ed with row & 7): byte offset of the four halves k = 4q .. 4q+3 of row r.  Same layout rule as the 128-row kernel's
// converter and as tc_prep_blob_kernel (halves: r*64 + (((k>>3) ^ (r&7)) << 3) + (k&7)).
TC64_HD uint32_t tc64_a_offset(int r, int q) {
  return (uint32_t)r * 128u + ((((uint32_t)q >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)q & 1u) * 8u;
}

// Accumulator layout of tcgen05.mma.cta_group::2 with M = 128 (64 rows per CTA), N = 256 -- cute's "2x2" atom
// (mma_traits_sm100.hpp, tmem_frg<.., N_SM = 2>, M_MMA_SM = 64): element (m, n) of this CTA's 64 x 256 tile lives in
// TMEM lane m + 64 * (n / 128), column n % 128.  An epilogue warp can read lanes [32 * quarter, +32) only (quarter = warp % 4)
// and scans columns [64 * sub, +64): the rows and codes it therefore sees are
TC64_HD int tc64_row_base(int quarter) { return (quarter & 1) * 32; }                       // + lane
TC64_HD int tc64_code_base(int quarter, int sub) { return (quarter >> 1) * 128 + sub * 64; }  // + column within the 64
TC64_HD int tc64_tmem_lane(int m, int n) { return m + 64 * (n >> 7); }
TC64_HD int tc64_tmem_col(int n) { return n & 127; }

#endif  // RQB200_TC64_LAYOUT_CUH
