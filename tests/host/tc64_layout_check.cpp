// Host check of csrc/tc64_layout.cuh (the index arithmetic of rq_tc64_kernel) against the layouts' definitions.
//   1. converter: every (row, k) of a 64 x 64 chunk is read from the fp32 staging stage exactly once and lands on the
//      half-precision position the K-major SWIZZLE_128B rule prescribes (the rule tc_prep_blob_kernel uses for B);
//      a warp-wide LDS.128 covers 512 contiguous bytes and a half-warp STS.64 one whole 128-byte row (no bank conflicts)
//   2. epilogue: the (TMEM lane, column) a thread reads equals cute's "2x2" accumulator layout for (its row, its code), and
//      the 8 warps cover the 64 x 256 score tile exactly once; the candidate-mask words of the 4 warps of a row tile [0,8)
#include "../../rq_vae_recommender_b200/csrc/tc64_layout.cuh"
#include <cstdio>
#include <vector>

int main() {
  int bad = 0;
  // ---- 1. converter
  std::vector<int> stage_hits(64 * 64, 0), a_hits(64 * 64, 0);
  for (int cw = 0; cw < 4; ++cw)
    for (int j = 0; j < 8; ++j) {
      uint32_t lo = ~0u, hi = 0;
      for (int lane = 0; lane < 32; ++lane) {
        const int h = lane >> 4, q = lane & 15, r = 16 * cw + 2 * j + h;
        const uint32_t so = tc64_stage_offset(r, q), ao = tc64_a_offset(r, q);
        if (so % 16 || ao % 8) ++bad;
        lo = so < lo ? so : lo; hi = so + 16 > hi ? so + 16 : hi;
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * q + e;
          ++stage_hits[(so + 4 * e) / 4];
          if ((int)(so + 4 * e) / 4 != r * 64 + k) ++bad;                       // row-major fp32 staging
          const int want = r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7);     // halves, SWIZZLE_128B K-major
          if ((int)(ao / 2) + e != want) ++bad;
          ++a_hits[ao / 2 + e];
        }
        if (ao / 128 != (uint32_t)r) ++bad;                                    // a half-warp stays inside its row
      }
      if (hi - lo != 512) ++bad;                                               // one LDS.128 = 512 contiguous bytes
    }
  for (int i = 0; i < 64 * 64; ++i) if (stage_hits[i] != 1 || a_hits[i] != 1) ++bad;
  // ---- 2. epilogue
  std::vector<int> tile_hits(64 * 256, 0), word_hits(64 * 8, 0);
  for (int quarter = 0; quarter < 4; ++quarter)
    for (int sub = 0; sub < 2; ++sub)
      for (int lane = 0; lane < 32; ++lane) {
        const int m = tc64_row_base(quarter) + lane, cb = tc64_code_base(quarter, sub);
        for (int c = 0; c < 64; ++c) {
          const int n = cb + c;
          if (tc64_tmem_lane(m, n) != quarter * 32 + lane) ++bad;
          if (tc64_tmem_col(n) != sub * 64 + c) ++bad;
          ++tile_hits[m * 256 + n];
        }
        for (int w = 0; w < 2; ++w) ++word_hits[m * 8 + (cb >> 5) + w];
      }
  for (int i = 0; i < 64 * 256; ++i) if (tile_hits[i] != 1) ++bad;
  for (int i = 0; i < 64 * 8; ++i) if (word_hits[i] != 1) ++bad;
  // kGrp = 2 (two epilogue groups on alternate tiles): a warp scans all 128 columns of its lane quarter; the 4 warps of a
  // group cover the tile once, the mask words of the 2 warps of a row tile [0,8)
  std::vector<int> tile2(64 * 256, 0), word2(64 * 8, 0);
  for (int quarter = 0; quarter < 4; ++quarter)
    for (int lane = 0; lane < 32; ++lane) {
      const int m = tc64_row_base(quarter) + lane, cb = (quarter >> 1) * 128;
      for (int c = 0; c < 128; ++c) {
        const int n = cb + c;
        if (tc64_tmem_lane(m, n) != quarter * 32 + lane || tc64_tmem_col(n) != c) ++bad;
        ++tile2[m * 256 + n];
      }
      for (int w = 0; w < 4; ++w) ++word2[m * 8 + (cb >> 5) + w];
    }
  for (int i = 0; i < 64 * 256; ++i) if (tile2[i] != 1) ++bad;
  for (int i = 0; i < 64 * 8; ++i) if (word2[i] != 1) ++bad;
  printf("tc64 layout check: bad %d \n", bad);
  return bad != 0;
}
