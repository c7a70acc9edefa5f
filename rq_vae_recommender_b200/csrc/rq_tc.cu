// Tensor-core tokeniser for sm_100a: tcgen05 fp16 candidate filter + exact fp32 re-rank.
//
// Result contract: identical to rqb200_rq_forward(mode = EVAL, ids only) -- the hard-argmin chain of
// modules/quantize.py:113-128,159-161 x L + modules/rqvae.py:125-132 (what semids.py:125 consumes).
//
// Why tensor cores: the distance term x.c^T is 2*D*K*L = 1.18 MFLOP per 3 KB item (381 FLOP/B);
// on CUDA cores the pass is ~30x compute bound.  Why it is still exact: the fp16 product only FILTERS.
//   S_l[b,k]  = fp16(x_b) . fp16(c_{l,k})            (tcgen05.mma, fp32 accumulate in TMEM; exact power-of-two scales)
//   score_l   = cc_{l,k} - 2 (S_l - sum_{j<l} G_{jl}[id_j, k])     (G = fp32 Gram tables C_j C_l^T, so every level is
//               scored from the ONE fp16 image of x: the residual never has to be re-quantised or re-staged)
//   candidates = { k : score <= min + 4 eps_b }      eps_b bounds the fp16 rounding of the dot product (margin in the epilogue)
//   |candidates| == 1  -> that code is the exact argmin;  else the candidates are re-scored with the exact fp32
//   arithmetic of the CUDA-core kernel (sequential fp32 residual, (xx + cc) - 2 dot, first index wins ties).
//
// Kernel structure (persistent, one CTA per SM, 128 rows per tile, 16 warps = 4 warpgroups, setmaxnreg 32 / 128 / 176):
//   WG0 warp 0   B producer  : 16 KB pre-swizzled fp16 codebook blocks -> 2-stage smem ring, TMA bulk copies + mbarriers
//   WG0 warp 1   MMA issuer  : one thread issues tcgen05.mma (M128 N128 K16), chunk-major (k chunk outer, column half
//                              inner), accumulators double-buffered in TMEM (2 x 256 columns)
//   WG1 (4 warps) converters : 128-bit coalesced fp32 loads of x -> fp16 -> K-major SWIZZLE_128B smem (A operand),
//                              refilled chunk by chunk as the last level releases it (x is read from HBM once);
//                              also the per-row max|x| and sum x^2 the filter margin needs
//   WG2-3 (8 warps) epilogue : 2 warps per TMEM lane quarter (one per 128-column half): tcgen05.ld scores, Gram
//                              correction, packed-key top-3 (tc_select.cuh), merge through smem, warp-cooperative
//                              exact re-rank from the fp32 codebook copy held in the prepared state
// Hot loops are kept SMALL on purpose (16-column rolled scan body, vector-only converter instantiation): the first
// versions were 100+ KB of straight-line SASS and ncu showed `no_inst` (instruction fetch) as the top stall of plain ALU
// instructions -- the L1.5 instruction cache is 32 KB, the per-scheduler L0 ~6 KB.
// With the 225 KB shared-memory carve-out there is no L1, so a register spill is an L2 round trip: the hot loops are kept
// spill-free (checked in SASS) and the role budgets sum to the CTA's launch allocation (setmaxnreg draws from it).
// Measured limits and the hypotheses tested on the way: DESIGN.md section 5.2, profiles/r1_tc_role_trace.txt.
#include "tc_common.cuh"

extern "C" int rqb200_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int64_t lda,
                            const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int relu,
                            const float* mask, int64_t ldmask, void* stream);


extern "C" int rqb200_tokenize_tc_supported(int D, int K, int L) {
  return (K == TC_K && D >= TC_KC && D <= TC_MAX_D && D % TC_KC == 0 && L >= 1 && L <= RQB_MAX_LEVELS) ? 1 : 0;
}

extern "C" size_t rqb200_tokenize_tc_state_bytes(int D, int K, int L) {
  if (!rqb200_tokenize_tc_supported(D, K, L)) return 0;
  return tc_off_blob(D, L) + (size_t)L * 2 * (D / TC_KC) * TC_BSTAGE_BYTES;
}

// ------------------------------------------------------------------------------------------------ prepare
__global__ void tc_prep_stats_kernel(const float* const* cbs, int D, TcHeader* hdr, float* cc) {
  const int l = blockIdx.y;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= TC_K) return;
  const float* c = cbs[l] + (int64_t)k * D;
  float s2 = 0.f, s4 = 0.f, s1 = 0.f, mx = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = c[d], v2 = v * v;
    s2 = fmaf(v, v, s2);
    s4 = fmaf(v2, v2, s4);
    s1 += fabsf(v);
    mx = fmaxf(mx, fabsf(v));
  }
  s2 = warp_sum(s2); s4 = warp_sum(s4); s1 = warp_sum(s1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    cc[l * TC_K + k] = s2;
    atomicMax(&hdr->amax_bits[l], __float_as_uint(mx));
    atomicMax(&hdr->c4_bits[l], __float_as_uint(sqrtf(s4)));
    atomicMax(&hdr->c1_bits[l], __float_as_uint(s1));
    atomicMax(&hdr->c2_bits[l], __float_as_uint(sqrtf(s2)));
  }
}

__global__ void tc_prep_consts_kernel(TcHeader* hdr, int L) {
  const int l = threadIdx.x;
  if (l >= L) return;
  const float amax = __uint_as_float(hdr->amax_bits[l]);
  float sc = 1.f;
  if (amax > 0.f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);          // amax = m * 2^e, m in [0.5, 1)
    e = max(-60, min(60, e));
    sc = ldexpf(1.f, -e);      // amax * sc in [0.5, 1)
  }
  TcLevelConst& c = hdr->lv[l];
  c.sc = sc;
  c.c4max = __uint_as_float(hdr->c4_bits[l]);
  c.c1max = __uint_as_float(hdr->c1_bits[l]);
  c.c2max = __uint_as_float(hdr->c2_bits[l]);
  float g = 0.f;
  for (int j = 0; j < l; ++j) g += __uint_as_float(hdr->c2_bits[j]);
  c.gerr = 3.8e-6f * g * c.c2max;   // 2^-18 * sum_j ||e_j|| ||c||: fp32 dot of length <= 768, generous
}

// Bblob[(l*2+h)*nkc + kc] = 16 KB smem image of codes [128h, 128h+128) x k [64kc, 64kc+64):
// K-major, 128 B per code row, 16-byte chunks XOR-swizzled with (row & 7)  (UMMA SWIZZLE_128B canonical layout)
__global__ void tc_prep_blob_kernel(const float* const* cbs, int D, const TcHeader* hdr, __half* blob) {
  const int nkc = D / TC_KC;
  const int blk = blockIdx.x;  // (l*2+h)*nkc + kc
  const int kc = blk % nkc, h = (blk / nkc) & 1, l = blk / (2 * nkc);
  const float sc = hdr->lv[l].sc;
  const float* c = cbs[l];
  __half* out = blob + (size_t)blk * (TC_BSTAGE_BYTES / 2);
  for (int i = threadIdx.x; i < 128 * TC_KC; i += blockDim.x) {
    const int n = i / TC_KC, k = i % TC_KC;
    const float v = c[(int64_t)(h * 128 + n) * D + kc * TC_KC + k] * sc;
    const int chunk = (k >> 3) ^ (n & 7);
    out[n * 64 + chunk * 8 + (k & 7)] = __float2half_rn(v);
  }
}

// hcc[l][k] = cc_l[k] / 2;  G_{0,l}[i][k] += cc_l[k] / 2  (l >= 1): the epilogue then scores with ONE table sum,
// half-distance h[k] = T[k] - S[k]*inv,  T = cc/2 + sum_j G_{j,l}[id_j]  (argmin-equivalent to quantize.py:113-117)
__global__ void tc_prep_fold_kernel(const float* cc, float* hcc, float* gram, int L) {
  const int l = blockIdx.y;
  const int k = threadIdx.x;   // 256 threads
  const float h = 0.5f * cc[l * TC_K + k];
  if (blockIdx.x == 0) hcc[l * TC_K + k] = h;
  if (l >= 1) {
    float* g = gram + (size_t)(l * (l - 1) / 2) * TC_K * TC_K;   // table (j = 0, l)
    for (int i = blockIdx.x; i < TC_K; i += gridDim.x) g[(size_t)i * TC_K + k] += h;
  }
}

extern "C" int rqb200_tokenize_tc_prepare(const float* const* codebooks, int D, int K, int L, void* state,
                                          size_t state_bytes, void* stream) {
  if (!rqb200_tokenize_tc_supported(D, K, L)) {
    rqb_set_error("tokenize_tc: shape D=%d K=%d L=%d not supported (need K=256, D %% 64 == 0, 64 <= D <= 768)", D, K, L);
    return RQB_ERR_UNSUPPORTED;
  }
  RQB_CHECK_ARG(codebooks && state, "tokenize_tc_prepare: null pointer");
  if (state_bytes < rqb200_tokenize_tc_state_bytes(D, K, L)) {
    rqb_set_error("tokenize_tc_prepare: state too small");
    return RQB_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* base = reinterpret_cast<char*>(state);
  TcHeader* hdr = reinterpret_cast<TcHeader*>(base);
  float* cc = reinterpret_cast<float*>(base + tc_off_cc(L));
  float* gram = reinterpret_cast<float*>(base + tc_off_gram(L));
  const float** cbptr = reinterpret_cast<const float**>(base + tc_off_cbptr(L));
  __half* blob = reinterpret_cast<__half*>(base + tc_off_blob(D, L));
  float* cbf = reinterpret_cast<float*>(base + tc_off_cbf(L));
  RQB_CUDA(cudaMemsetAsync(hdr, 0, sizeof(TcHeader), st));
  RQB_CUDA(cudaMemcpyAsync(cbptr, codebooks, sizeof(float*) * L, cudaMemcpyHostToDevice, st));
  // fp32 copy for the exact re-rank: 256-byte aligned rows whatever the caller's tensors look like, and the prepared state
  // no longer references caller memory after this call returns (stream order)
  for (int l = 0; l < L; ++l)
    RQB_CUDA(cudaMemcpyAsync(cbf + (size_t)l * TC_K * D, codebooks[l], sizeof(float) * TC_K * D, cudaMemcpyDeviceToDevice, st));
  tc_prep_stats_kernel<<<dim3(TC_K / 8, L), 256, 0, st>>>(cbptr, D, hdr, cc);
  RQB_LAUNCH_CHECK();
  tc_prep_consts_kernel<<<1, 32, 0, st>>>(hdr, L);
  RQB_LAUNCH_CHECK();
  tc_prep_blob_kernel<<<L * 2 * (D / TC_KC), 256, 0, st>>>(cbptr, D, hdr, blob);
  RQB_LAUNCH_CHECK();
  for (int l = 1; l < L; ++l)
    for (int j = 0; j < l; ++j) {
      float* g = gram + (size_t)(l * (l - 1) / 2 + j) * TC_K * TC_K;
      int rc = rqb200_sgemm(0, 1, TC_K, TC_K, D, 1.f, codebooks[j], D, codebooks[l], D, 0.f, g, TC_K, 0, nullptr, 0, stream);
      if (rc) return rc;
    }
  float* hcc = reinterpret_cast<float*>(base + tc_off_hcc(L));
  tc_prep_fold_kernel<<<dim3(32, L), TC_K, 0, st>>>(cc, hcc, gram, L);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}


struct TcSmemMisc {
  uint64_t a_full[TC_MAX_KC], a_empty[TC_MAX_KC];
  uint64_t b_full[TC_BSTAGES], b_empty[TC_BSTAGES];
  uint64_t t_full[2][2];          // [accumulator buffer][column half]
  uint64_t t_empty[2];
  uint64_t rowinfo_free;
  uint32_t tmem_base;
  uint32_t pad;
  uint32_t rowinfo[TC_BM];        // bf16x2 (rounded up): max|x| | sum x^2 of the tile being scored
  TcExch exch[TC_BM];             // half-1 warp -> half-0 warp of the same lane quarter
  uint64_t xs_full, xs_free;      // kTma only: fp32 staging of x landed in the A slots / read out by every converter thread
  uint32_t conv_sink, conv_pad;   // kTma only: dependency sink of the converters (see tc_conv_sync_after)
};


// kTrace = true compiles the clock64 role accounting in (RQB200_TC_TRACE=1); the production instantiation carries none of it.
// kVec = false is the slow-path instantiation for x whose rows are not 16-byte aligned (scalar loads); keeping it out of the
// main instantiation halves the converter's code.
// kPair = true is the CTA-pair instantiation (launched as clusters of 2, tcgen05 cta_group::2): the two CTAs of a pair take the
// two 128-row halves of a 256-row pair-tile; ONE M=256 x N=256 instruction of the leader drives both tensor cores, each CTA
// streams only ITS half of every codebook block (half the L2 and shared-memory traffic per row, and the same 2 x 16 KB ring
// now covers twice the tensor time).  Barriers the leader waits on (a_full, b_full, t_empty) collect arrivals from both CTAs;
// barriers the leader signals (a_empty, b_empty, t_full) are multicast commits.
// kOpt bit 0 = kTma (RQB200_TC_TMA=1, opt-in; first hardware run raced, fixed since, fix unrun): x reaches the converter through TMA instead of the LSU path,
// staged IN PLACE in the A slots the chunk is about to occupy -- there is no other shared memory left in this kernel.  The
// fp32 source of chunk kc is two 16 KB boxes (128 rows x 32 floats): box 0 lands in slot kc, box 1 in slot kc + 1 (the next
// chunk's slot, already released by the previous tile); the converters pull both into registers, meet at a named barrier,
// and only then write the fp16 image into slot kc, while warp 2 already fetches the next chunk into slots kc + 1 / kc + 2.
// The last chunk has no next slot: its two boxes go through its own slot one after the other.  One chunk (32 KB) is in
// flight at a time, but as bulk copies: no L1TEX miss tracking, no LSU queue shared with the epilogue's gathers (the measured
// limit of the register path is 5-6 B/clk/SM at a ~6 K-cycle loaded latency).
// kOpt bit 1 (RQB200_TC_FASTSCAN=1, opt-in, not yet run): the scan arithmetic of rq_tc64_kernel -- two scores per FFMA2 and per
// top-3 insertion, one-instruction key packs, FADD2 Gram folds (about 14 -> 10 SASS instructions per score).
template <bool kTrace, bool kVec, bool kPair, int kOpt = 0>
__global__ void __launch_bounds__(TC_THREADS, 1) rq_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr bool kTma = (kOpt & 1) != 0, kFast = (kOpt & 2) != 0;
  extern __shared__ __align__(1024) unsigned char tsm[];
  unsigned char* sA = tsm;                                         // [nkc][16 KB]  (sized for TC_MAX_KC)
  unsigned char* sB = tsm + TC_MAX_KC * TC_ACHUNK_BYTES;           // [TC_BSTAGES][16 KB]
  TcSmemMisc* ms = reinterpret_cast<TcSmemMisc*>(sB + TC_BSTAGES * TC_BSTAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkc = p.nkc, L = p.L;
  const bool trace = kTrace && p.stats != nullptr;
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;              // 0 = leader
  // work units: tiles (one CTA each) or pair-tiles (one cluster each; the CTA takes tile 2*unit + crank, which may lie
  // past the last tile when the tile count is odd: that CTA still runs the whole protocol on clamped rows and stores nothing)
  const int u_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int u_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int u_count = kPair ? (p.ntiles + 1) >> 1 : p.ntiles;
  #define TC_TILE_OF(unit) (kPair ? 2 * (unit) + (int)crank : (unit))
  const int rot0 = p.rot ? u_first % nkc : 0;     // both CTAs of a pair must walk the chunks in the same order

  if (tid == 0) {
    if ((smem_u32(tsm) & 1023u) != 0) __trap();  // the swizzle pattern needs a 1024-byte aligned base
    // pair: one aggregated arrive per converter / epilogue warp of EACH CTA (remote arrives are per warp, not per thread)
    for (int i = 0; i < TC_MAX_KC; ++i) { mbar_init(&ms->a_full[i], kPair ? 2 * TC_NCONV_WARPS : TC_NCONV_WARPS * 32); mbar_init(&ms->a_empty[i], 1); }
    for (int i = 0; i < TC_BSTAGES; ++i) { mbar_init(&ms->b_full[i], 1); mbar_init(&ms->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ms->t_full[i][0], 1);
      mbar_init(&ms->t_full[i][1], 1);
      mbar_init(&ms->t_empty[i], kPair ? 2 * TC_NEPI_WARPS : TC_NEPI_WARPS * 32);
    }
    mbar_init(&ms->rowinfo_free, 128);   // the half-0 epilogue thread of every row
    if (kTma) { mbar_init(&ms->xs_full, 1); mbar_init(&ms->xs_free, 1); }
    fence_mbar_init();
  }
  if (warp == 1) { if (kPair) tc_alloc2(&ms->tmem_base, 512); else tc_alloc(&ms->tmem_base, 512); }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();     // the peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  #define TC_TMEM_BASE() (*reinterpret_cast<volatile uint32_t*>(&ms->tmem_base))

  if (warp < 4) {
    // ============================================================== warpgroup 0: B producer (warp 0), MMA issuer (warp 1)
    // register budget (per-CTA pool = 512 threads x 128 at launch = 65536): 128x32 + 128x128 + 256x176 = 65536.
    // With the 225 KB shared-memory carve-out there is no L1: a spill is an L2 round trip, so no role may spill in a loop.
    tc_setmaxnreg_dec<32>();
    if (warp == 0) {
      uint32_t s = 0;
      for (int unit = u_first; unit < u_count; unit += u_step) {
        if (kVec && p.prefetch && lane == 0) {
          // pull the NEXT tile's x rows into L2 now (393 KB per CTA, 58 MB chip-wide: fits the 126 MB L2), so the converter's
          // loads are L2 hits: with 32 KB of register buffers in flight per SM it cannot cover the HBM latency-bandwidth
          // product (23 B/clk x ~1500 cycles), it can cover L2's
          const int nt = TC_TILE_OF(unit + u_step);
          if (unit + u_step < u_count && nt < p.ntiles) {
            const int rows = min(TC_BM, p.B - nt * TC_BM);
            const float* xn = p.x + (int64_t)nt * TC_BM * p.ldx;
#pragma unroll 1
            for (int r = 0; r < rows; ++r) bulk_prefetch_l2(xn + (int64_t)r * p.ldx, (uint32_t)p.D * 4u);
          }
        }
        if constexpr (kPair) {
          // one 16 KB block per chunk step: THIS CTA's 128 codes (column half = crank) of level l, chunk kc; the bytes of both
          // CTAs are counted on the LEADER's b_full, which is what its MMA warp waits on
          for (int l = 0; l < L; ++l)
            for (int i = 0; i < nkc; ++i, ++s) {
              const int kc = tc_rot(i, rot0, nkc);
              const uint32_t st = s % TC_BSTAGES, u = s / TC_BSTAGES;
              mbar_wait_guarded(&ms->b_empty[st], (u & 1) ^ 1, 1);      // local: the leader's commits are multicast
              if (tc_elect_one()) {
                if (crank == 0) mbar_expect_tx(&ms->b_full[st], 2 * TC_BSTAGE_BYTES);
                tc_tma2d_pair(sB + st * TC_BSTAGE_BYTES, &p.tmapB, 0, ((l * 2 + (int)crank) * nkc + kc) * 128,
                              cluster_map(smem_u32(&ms->b_full[st]), 0));
              }
              __syncwarp();
            }
        } else
        for (int l = 0; l < L; ++l)
          for (int i = 0; i < nkc; ++i)
            for (int h = 0; h < 2; ++h, ++s) {      // same order as the MMA issuer: chunk-major, column half inner
              const int kc = tc_rot(i, rot0, nkc);
              const uint32_t st = s % TC_BSTAGES, u = s / TC_BSTAGES;
              mbar_wait_guarded(&ms->b_empty[st], (u & 1) ^ 1, 1);
              if (tc_elect_one()) {
                mbar_expect_tx(&ms->b_full[st], TC_BSTAGE_BYTES);
                bulk_g2s(sB + st * TC_BSTAGE_BYTES, p.blob + (size_t)((l * 2 + h) * nkc + kc) * TC_BSTAGE_BYTES,
                         TC_BSTAGE_BYTES, &ms->b_full[st]);
              }
              __syncwarp();
            }
      }
    } else if (warp == 1 && crank == 0) {
      const uint32_t idesc = kPair ? tc_idesc(256, 256) : tc_idesc(128, 128);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      uint32_t s = 0, g = 0, it = 0;
      long long w_te = 0, w_af = 0, w_bf = 0, w_issue = 0;
      TC_EV_DECL();
      TC_T0(tm);
      const long long tm_start = tm;
      for (int unit = u_first; unit < u_count; unit += u_step, ++it)
        for (int l = 0; l < L; ++l, ++g) {
          const uint32_t buf = g & 1, u = g >> 1;
          TC_ACC(w_issue, tm);
          if (kPair) mbar_wait_guarded_cluster(&ms->t_empty[buf], (u & 1) ^ 1, 2);
          else mbar_wait_guarded(&ms->t_empty[buf], (u & 1) ^ 1, 2);
          TC_ACC(w_te, tm);
          TC_EV(0, 1, it * 16 + l);
          tc_fence_after();
          // chunk-major order (kc outer, column half inner): both 128-column halves of the score tile complete together, so
          // neither epilogue warp of a lane quarter waits for the other's scan to start, and at the last level chunk kc is
          // released after 2 steps instead of 12 + kc, which widens the window for refilling A with the next tile.
          const uint32_t d_base = TC_TMEM_BASE() + buf * 256;
          for (int i = 0; i < nkc; ++i) {
            const int kc = tc_rot(i, rot0, nkc);
            TC_ACC(w_issue, tm);
            if (l == 0) {
              if (kPair) mbar_wait_guarded_cluster(&ms->a_full[kc], it & 1, 3);
              else mbar_wait_guarded(&ms->a_full[kc], it & 1, 3);
            }
            TC_ACC(w_af, tm);
            if (l == 0) TC_EV(0, 2, it * 16 + i);
            const uint64_t adesc = tc_smem_desc(a_base + kc * TC_ACHUNK_BYTES);
            if constexpr (kPair) {
              const uint32_t st = s % TC_BSTAGES;
              mbar_wait_guarded_cluster(&ms->b_full[st], (s / TC_BSTAGES) & 1, 4);
              TC_ACC(w_bf, tm);
              tc_fence_after();
              const uint64_t bdesc = tc_smem_desc(b_base + st * TC_BSTAGE_BYTES);
              if (tc_elect_one()) {
#pragma unroll
                for (int j = 0; j < TC_KC / 16; ++j)
                  tc_mma_f16_2(d_base, adesc + 2 * j, bdesc + 2 * j, idesc, (i | j) != 0);
                tc_commit2(&ms->b_empty[st]);
                if (l == L - 1) tc_commit2(&ms->a_empty[kc]);
                if (i == nkc - 1) {
                  tc_commit2(&ms->t_full[buf][0]);
                  tc_commit2(&ms->t_full[buf][1]);
                }
              }
              __syncwarp();
              ++s;
            } else
            for (int h = 0; h < 2; ++h, ++s) {
              const uint32_t st = s % TC_BSTAGES;
              mbar_wait_guarded(&ms->b_full[st], (s / TC_BSTAGES) & 1, 4);
              TC_ACC(w_bf, tm);
              tc_fence_after();
              const uint64_t bdesc = tc_smem_desc(b_base + st * TC_BSTAGE_BYTES);
              if (tc_elect_one()) {
#pragma unroll
                for (int j = 0; j < TC_KC / 16; ++j)   // K=16 per instruction: +32 B inside the 128 B swizzle row
                  tc_mma_f16(d_base + h * 128, adesc + 2 * j, bdesc + 2 * j, idesc, (i | j) != 0);
                tc_commit(&ms->b_empty[st]);
                if (h == 1 && l == L - 1) tc_commit(&ms->a_empty[kc]);
                if (h == 1 && i == nkc - 1) {
                  tc_commit(&ms->t_full[buf][0]);
                  tc_commit(&ms->t_full[buf][1]);
                }
              }
              __syncwarp();
            }
          }
          TC_EV(0, 3, it * 16 + l);
        }
      if (trace && lane == 0) {
        tc_trace_add(p.stats, 0, w_te); tc_trace_add(p.stats, 1, w_af); tc_trace_add(p.stats, 2, w_bf);
        tc_trace_add(p.stats, 3, clock64() - tm_start); tc_trace_add(p.stats, 12, 1);
      }
    } else if (kTma && warp == 2) {
      // x producer of the in-place TMA path: one load step = the fp32 boxes of one chunk (two boxes, or one at a time for the
      // last chunk of a tile).  A step may start when (a) the slots it lands in were released by the previous tile's last
      // level and (b) every converter thread has read the previous step's boxes out (its box 1 sat in this step's slot).
      uint32_t ls = 0, it = 0;
      for (int unit = u_first; unit < u_count; unit += u_step, ++it) {
        const int tile = min(TC_TILE_OF(unit), p.ntiles - 1);   // a pair's second CTA past the last tile loads that tile again (unused)
        const int row0 = tile * TC_BM;                           // rows past B read as zero (tensor-map bounds)
        for (int kc = 0; kc < nkc; ++kc) {
          const bool last_chunk = (kc == nkc - 1);
          mbar_wait_guarded(&ms->a_empty[kc], (it & 1) ^ 1, 13);
          if (!last_chunk) mbar_wait_guarded(&ms->a_empty[kc + 1], (it & 1) ^ 1, 13);
          mbar_wait_guarded(&ms->xs_free, (ls & 1) ^ 1, 14);
          if (tc_elect_one()) {
            mbar_expect_tx(&ms->xs_full, last_chunk ? TC_ACHUNK_BYTES : 2 * TC_ACHUNK_BYTES);
            tc_tma2d(sA + kc * TC_ACHUNK_BYTES, &p.tmapXh, kc * TC_KC, row0, &ms->xs_full);
            if (!last_chunk) tc_tma2d(sA + (kc + 1) * TC_ACHUNK_BYTES, &p.tmapXh, kc * TC_KC + 32, row0, &ms->xs_full);
          }
          __syncwarp();
          ++ls;
          if (last_chunk) {
            mbar_wait_guarded(&ms->xs_free, (ls & 1) ^ 1, 14);
            if (tc_elect_one()) {
              mbar_expect_tx(&ms->xs_full, TC_ACHUNK_BYTES);
              tc_tma2d(sA + kc * TC_ACHUNK_BYTES, &p.tmapXh, kc * TC_KC + 32, row0, &ms->xs_full);
            }
            __syncwarp();
            ++ls;
          }
        }
      }
    }
  } else if (warp < 4 + TC_NCONV_WARPS) {
    // ============================================================== warpgroup 1: x fp32 -> fp16 swizzled A chunks
    // Warp cw owns rows [32cw, 32cw+32) of every chunk.  Lane -> row 4i + (lane >> 3), i = 0..7 (8 rows per lane, so the
    // row statistics are 16 registers) and float4 column (lane & 7) + 8j: half-unit j = 0/1 is 8 x LDG.128 per lane,
    // each warp instruction reading four 128-byte segments.  The two register buffers leapfrog, so 8..16 loads per
    // lane are always in flight and nothing is copied.
    // (stays at the launch allocation of 128 registers: at 96 this loop spilled its row pointers -> L2 latency per chunk)
    const int cw = warp - 4;
    const int rsub = lane >> 3, q8 = lane & 7;
    uint32_t it = 0;
    [[maybe_unused]] uint32_t ls = 0;     // kTma: load steps consumed (phase of xs_full)
    long long c_wait = 0, c_work = 0, c_ldwait = 0, c_cvt = 0;
    TC_EV_DECL();
    TC_T0(tcv);
    const long long tcv_start = tcv;
#pragma unroll 1
    for (int unit = u_first; unit < u_count; unit += u_step, ++it) {
      const int tile = min(TC_TILE_OF(unit), p.ntiles - 1);   // a pair's second CTA past the last tile converts that tile again (unused)
      // row statistics for the filter margin: max|x| and sum x^2 (sum x^4 <= max|x|^2 * sum x^2 is used downstream; max|x|
      // doubles as the fp16 overflow test, NaN inputs surface through sum x^2).  ~half the ALU of tracking sum x^4 and
      // testing every converted half for inf.
      float sm[8], s2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { sm[i] = 0.f; s2[i] = 0.f; }
      float4 va[8], vb[8];
      // half-unit h: chunk kc = h >> 1, float4 columns 8*(h & 1) + q8 (= element 32 h + 4 q8 of the row), rows row_base + 4i.
      // Register- and instruction-frugal on purpose: ONE 64-bit tile pointer plus eight 32-bit row offsets.  Rows past B
      // are CLAMPED to row B-1 instead of predicated: no zero-fill, no per-load compare, one code path; their scores are
      // never stored.  (host side guarantees 127 * ldx < 2^32)
      const float* xt = p.x + (int64_t)tile * TC_BM * p.ldx + q8 * 4;
      const int last = p.B - 1 - tile * TC_BM;                           // >= 0: last valid row of this tile
      uint32_t off[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) off[i] = (uint32_t)min(cw * 32 + rsub + 4 * i, last) * (uint32_t)p.ldx;
      auto load_half = [&](float4 (&v)[8], int h) {
        const float* xh = xt + h * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* src = xh + off[i];
          if (kVec) v[i] = ldg_stream(reinterpret_cast<const float4*>(src));
          else { v[i].x = __ldg(src); v[i].y = __ldg(src + 1); v[i].z = __ldg(src + 2); v[i].w = __ldg(src + 3); }
        }
      };
      const uint32_t srow = (uint32_t)(cw * 32 + rsub) * 128;            // byte offset of row (cw*32 + rsub) in a chunk
      auto convert_half = [&](const float4 (&v)[8], int h) {
        const uint32_t f = q8 + 8 * (h & 1);
        // 16-byte chunk position after the 128B swizzle, for even / odd i (row & 7 = rsub or rsub + 4)
        const uint32_t base_e = smem_u32(sA) + (h >> 1) * TC_ACHUNK_BYTES + srow + (((f >> 1) ^ (uint32_t)rsub) << 4) + (f & 1) * 8;
        const uint32_t base_o = base_e ^ (4u << 4);                      // (c ^ (rsub + 4)) = (c ^ rsub) ^ 4 since rsub < 4
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = v[i];
          s2[i] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, s2[i]))));
          sm[i] = fmaxf(fmaxf(sm[i], fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
          const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
          const uint32_t addr = ((i & 1) ? base_o : base_e) + i * 512;   // rows advance by 4 -> 4 * 128 bytes
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h0)),
                       "r"(*reinterpret_cast<const uint32_t*>(&h1)) : "memory");
        }
      };
      if constexpr (kTma) {
        // in-place TMA path (see the kernel's header comment): box 0 of chunk kc sits in slot kc, box 1 in slot kc + 1 (or, for
        // the last chunk, follows box 0 through slot kc).  Same lane -> (row, float4 column) mapping as the register path, so
        // convert_half is shared: LDS.128 of a warp covers four whole 128-byte staging rows (conflict-free).
        auto lds_half = [&](float4 (&v)[8], int slot) {
          const unsigned char* sp = sA + slot * TC_ACHUNK_BYTES + srow + q8 * 16;
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(sp + i * 512);
        };
        // convert_half split in two: statistics + fp16 packing BEFORE the converters' barrier, the stores after it.  The
        // packing consumes the loaded registers, so every LDS has returned its data before the barrier is signalled: the
        // barrier alone only orders the ISSUE of the loads, and the bulk copy that refills the slot does not pass through
        // the LSU queue they may still be waiting in (first hardware run: the un-split version returned a few wrong ids).
        auto pack_half = [&](const float4 (&v)[8], uint2 (&hp)[8]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = v[i];
            s2[i] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, s2[i]))));
            sm[i] = fmaxf(fmaxf(sm[i], fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
            const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
            hp[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
          }
        };
        auto store_half = [&](const uint2 (&hp)[8], int h) {     // same addresses as convert_half
          const uint32_t f = q8 + 8 * (h & 1);
          const uint32_t base_e = smem_u32(sA) + (h >> 1) * TC_ACHUNK_BYTES + srow + (((f >> 1) ^ (uint32_t)rsub) << 4) + (f & 1) * 8;
          const uint32_t base_o = base_e ^ (4u << 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t addr = ((i & 1) ? base_o : base_e) + i * 512;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(hp[i].x), "r"(hp[i].y) : "memory");
          }
        };
        // Before the barrier every thread STORES a value computed from every packed register (a real side effect: an unused asm
        // operand is dropped, and without it ptxas sinks most of the packing below the barrier -- seen in SASS -- so that the
        // loads are still in flight there).
        auto dep_of = [&](const uint2 (&hp)[8]) {
          uint32_t d = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) d ^= hp[i].x ^ hp[i].y;
          return d;
        };
        uint2 pa[8], pb[8];
#pragma unroll 1
        for (int kc = 0; kc < nkc; ++kc) {
          const bool last_chunk = (kc == nkc - 1);
          mbar_wait_guarded(&ms->xs_full, ls & 1, 12);
          if (cw == 0) TC_EV(1, 1, it * 16 + kc);
          lds_half(va, kc);
          if (!last_chunk) lds_half(vb, kc + 1);
          pack_half(va, pa);
          uint32_t dep = dep_of(pa);
          if (!last_chunk) { pack_half(vb, pb); dep ^= dep_of(pb); }
          tc_conv_sync_after(dep, &ms->conv_sink);   // every converter thread HOLDS its staging bytes: the slots may be overwritten
          if (cw == 0 && lane == 0) mbar_arrive(&ms->xs_free);
          ++ls;
          if (last_chunk) {
            mbar_wait_guarded(&ms->xs_full, ls & 1, 12);
            lds_half(vb, kc);
            pack_half(vb, pb);
            tc_conv_sync_after(dep_of(pb), &ms->conv_sink);
            if (cw == 0 && lane == 0) mbar_arrive(&ms->xs_free);
            ++ls;
          }
          store_half(pa, 2 * kc);
          store_half(pb, 2 * kc + 1);
          if (last_chunk) {
            mbar_wait_guarded(&ms->rowinfo_free, (it & 1) ^ 1, 6);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
              for (int o = 4; o > 0; o >>= 1) {      // the 8 lanes (q8) that share row 4r + rsub
                sm[r] = fmaxf(sm[r], __shfl_xor_sync(0xffffffffu, sm[r], o));
                s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
              }
              if (q8 == 0) ms->rowinfo[cw * 32 + rsub + 4 * r] = (tc_bf16_up(sm[r]) << 16) | tc_bf16_up(s2[r]);
            }
          }
          fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor-core (async) proxy
          if constexpr (kPair) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(cluster_map(smem_u32(&ms->a_full[kc]), 0));
          } else {
            mbar_arrive(&ms->a_full[kc]);
          }
          if (cw == 0) TC_EV(1, 2, it * 16 + kc);
        }
        continue;                              // next tile
      }
      // chunks are produced in the order the MMA issuer consumes them: step i -> chunk tc_rot(i)
      load_half(va, 2 * rot0);
      load_half(vb, 2 * rot0 + 1);
#pragma unroll 1
      for (int i = 0; i < nkc; ++i) {
        const int kc = tc_rot(i, rot0, nkc);
        const int kn = tc_rot(i + 1 < nkc ? i + 1 : i, rot0, nkc);    // next chunk (the last step re-loads its own: unused)
        TC_ACC(c_work, tcv);
        mbar_wait_guarded(&ms->a_empty[kc], (it & 1) ^ 1, 5);   // the last level of the previous tile released this chunk
        TC_ACC(c_wait, tcv);
        if (cw == 0) TC_EV(1, 1, it * 16 + i);
        if (trace) {   // split "waiting for the loads to land" from "convert + store": shfl needs the last-issued value
          const float probe = __shfl_sync(0xffffffffu, va[7].w, 0);
          asm volatile("" ::"f"(probe));
          TC_ACC(c_ldwait, tcv);
        }
        convert_half(va, 2 * kc);
        TC_ACC(c_cvt, tcv);
        if (i + 1 < nkc) load_half(va, 2 * kn);
        if (trace) {
          const float probe = __shfl_sync(0xffffffffu, vb[7].w, 0);
          asm volatile("" ::"f"(probe));
          TC_ACC(c_ldwait, tcv);
        }
        convert_half(vb, 2 * kc + 1);
        TC_ACC(c_cvt, tcv);
        if (i + 1 < nkc) load_half(vb, 2 * kn + 1);
        if (i == nkc - 1) {
          // row statistics for the margin: reduce over the 16 lanes that share a row, publish before the last arrive
          mbar_wait_guarded(&ms->rowinfo_free, (it & 1) ^ 1, 6);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {      // the 8 lanes (q8) that share row 4r + rsub
              sm[r] = fmaxf(sm[r], __shfl_xor_sync(0xffffffffu, sm[r], o));
              s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
            }
            if (q8 == 0) ms->rowinfo[cw * 32 + rsub + 4 * r] = (tc_bf16_up(sm[r]) << 16) | tc_bf16_up(s2[r]);
          }
        }
        fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        if constexpr (kPair) {               // one arrive per warp on the LEADER's barrier (remote for the peer CTA)
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(cluster_map(smem_u32(&ms->a_full[kc]), 0));
        } else {
          mbar_arrive(&ms->a_full[kc]);
        }
        if (cw == 0) TC_EV(1, 2, it * 16 + i);
      }
    }
    if (trace && cw == 0 && lane == 0) {
      tc_trace_add(p.stats, 9, c_wait); tc_trace_add(p.stats, 10, clock64() - tcv_start);
      tc_trace_add(p.stats, 20, c_ldwait); tc_trace_add(p.stats, 21, c_cvt);
    }
  } else {
    // ============================================================== warpgroups 3-4: scores -> candidates -> exact re-rank -> ids
    // two warps per TMEM lane quarter: `half` 0 scans columns [0,128) and owns merge / re-rank / ids, half 1 scans [128,256)
    tc_setmaxnreg_inc<176>();
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = (warp - (4 + TC_NCONV_WARPS)) >> 2;
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int bar_x = 2 + quarter;      // half 1 -> half 0: exch[] written
    const int bar_i = 6 + quarter;      // half 0 -> half 1: the level's id is final (written into exch[].idx)
    const int D = p.D;
    const int lane4 = lane * 4;
    uint32_t g = 0, it = 0;
    long long e_tf = 0, e_scan = 0, e_pair = 0, e_rr = 0, e_idw = 0, e_merge = 0, e_many = 0;
    TC_EV_DECL();
    const int ev_role = 2 + half;      // lane quarter 0 only
    TC_T0(te);
    const long long te_start = te;
    // accumulator buffer released: per thread locally, or one arrive per warp on the LEADER's barrier in the pair variant
    auto release_tmem = [&](uint32_t buf) {
      tc_fence_before();
      if constexpr (kPair) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(cluster_map(smem_u32(&ms->t_empty[buf]), 0));
      } else {
        mbar_arrive(&ms->t_empty[buf]);
      }
    };
#pragma unroll 1
    for (int unit = u_first; unit < u_count; unit += u_step, ++it) {
      const int tile = TC_TILE_OF(unit);
      const int row = tile * TC_BM + r_local;
      const bool valid = row < p.B;    // rows past B run the same code on zero scores (no divergent copies); nothing of theirs is stored
      uint64_t idpack = 0;          // 8 bits per level
      float x4s = 0.f, x2s = 0.f;
#pragma unroll 1
      for (int l = 0; l < L; ++l, ++g) {
        const uint32_t buf = g & 1, u = g >> 1;
        // T rows (full 256-column rows): hcc at level 0, else the Gram rows of the codes chosen at levels j < l, with cc/2
        // folded into table j = 0.  Two rows are pipelined statically (L <= 3 is the fast path); rows j >= 2 are summed in.
        const int tri = l * (l - 1) / 2;
        auto grow = [&](int j) -> const float* {
          return p.gram + ((size_t)(tri + j) * TC_K + (size_t)((idpack >> (8 * j)) & 0xff)) * TC_K;
        };
        const float* trow0 = (l >= 1) ? grow(0) : p.hcc;
        const float* trow1 = (l >= 2) ? grow(1) : trow0;
        auto load_t = [&](float4 (&ta)[4], float4 (&tb)[4], int col) {      // issue only: nothing here waits for data
#pragma unroll
          for (int v = 0; v < 4; v += 2) ldg256_pinned(trow0 + col + 4 * v, ta[v], ta[v + 1]);
          if (l >= 2) {
#pragma unroll
            for (int v = 0; v < 4; v += 2) ldg256_pinned(trow1 + col + 4 * v, tb[v], tb[v + 1]);
          }
        };
        auto fold_t = [&](float4 (&ta)[4], const float4 (&tb)[4], int col) {   // called one chunk of compute after load_t
          if (l >= 2) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if constexpr (kFast) { tc_add2(ta[v].x, ta[v].y, tb[v].x, tb[v].y); tc_add2(ta[v].z, ta[v].w, tb[v].z, tb[v].w); }
              else { ta[v].x += tb[v].x; ta[v].y += tb[v].y; ta[v].z += tb[v].z; ta[v].w += tb[v].w; }
            }
#pragma unroll 1
            for (int j = 2; j < l; ++j) {   // L > 3 only: latency exposed, code kept small
              const float* gj = grow(j) + col;
#pragma unroll
              for (int v = 0; v < 4; v += 2) {
                float4 t0, t1;
                ldg256_pinned(gj + 4 * v, t0, t1);
                ta[v].x += t0.x; ta[v].y += t0.y; ta[v].z += t0.z; ta[v].w += t0.w;
                ta[v + 1].x += t1.x; ta[v + 1].y += t1.y; ta[v + 1].z += t1.z; ta[v + 1].w += t1.w;
              }
            }
          }
        };
        const int col0 = half * 128;
        float4 ta0[4], tb0[4], ta1[4], tb1[4];
        load_t(ta0, tb0, col0);          // in flight across the accumulator wait below
        TC_ACC(e_rr, te);
        mbar_wait_guarded(&ms->t_full[buf][half], u & 1, 7);
        TC_ACC(e_tf, te);
        if (quarter == 0) TC_EV(ev_role, 1, it * 16 + l);
        tc_fence_after();
        const uint32_t tcol = TC_TMEM_BASE() + lane_addr + buf * 256 + col0;
        uint32_t s0[16], s1[16];
        tc_ld16_issue(tcol, s0);
        const TcLevelConst lc = p.hdr->lv[l];
        if (l == 0 && half == 0) {       // only the merging warp needs the margin
          const uint32_t ri = ms->rowinfo[r_local];
          const float xmax = __uint_as_float(ri & 0xffff0000u);        // max|x| (bf16, rounded up)
          x2s = __uint_as_float(ri << 16);                              // sum x^2 (bf16, rounded up; NaN if any input is)
          x4s = (xmax * p.sx < 65504.f) ? xmax * xmax * x2s : INFINITY;  // sum x^4 <= max|x|^2 sum x^2; fp16 overflow/inf -> poison
          mbar_arrive(&ms->rowinfo_free);
        }
        // ---- margin (DESIGN.md "filter error bound"): eps bounds |approx dot - exact dot|; scores are half-distances
        const float x2n = sqrtf(x2s);
        const float sig = 4.8828125e-4f * 0.81649658f * sqrtf(sqrtf(x4s) * lc.c4max);          // u=2^-11, sqrt(2/3)
        const float flo = 2.98023224e-8f * (lc.c1max / p.sx + sqrtf((float)p.D) * x2n / lc.sc);  // fp16 subnormal floor
        const float acc = 7.62939453e-6f * x2n * lc.c2max;                                      // 64 * 2^-23 accumulate
        const float eps = TC_Z * sig + flo + acc + lc.gerr;
        const float margin = 2.f * eps;
        const float ninv = -1.f / (p.sx * lc.sc);

        float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
        int i1 = 0, i2 = 0;
        // one 16-column chunk: packed-key triple (7 instructions per score), then one merge into the running top-3
        auto score16 = [&](const uint32_t (&s)[16], const float4 (&t)[4], int col) {
          // The always-true opaque branch makes this ALU block its own basic block.  ptxas otherwise treats the prefetches
          // issued just above it (Gram rows: an L2 round trip, ~700 cycles; next TMEM columns) as ordinary short loads and
          // sinks them to the END of the block, a few instructions ahead of their first use -- the software pipeline then
          // hides nothing (ncu: 660 long-scoreboard samples on the first FFMA after the loads).  Volatile asm and empty-asm
          // pins fix the order in PTX but not in SASS; a block boundary does.
          if (p.one) {
            float q1 = INFINITY, q2 = INFINITY, q3 = INFINITY;
            if constexpr (kFast) {
              const uint32_t kmask = p.one ? TCS_KEY_MASK : 0u;      // the key mask in a REGISTER (tcs_pack_reg)
#define TC_SCORE4(V)                                                                                                     \
              {                                                                                                            \
                float h0, h1, h2, h3;                                                                                      \
                tc_fma2(h0, h1, __uint_as_float(s[(V) * 4 + 0]), __uint_as_float(s[(V) * 4 + 1]), ninv, t[V].x, t[V].y);  \
                tc_fma2(h2, h3, __uint_as_float(s[(V) * 4 + 2]), __uint_as_float(s[(V) * 4 + 3]), ninv, t[V].z, t[V].w);  \
                tcs_key_insert2_keys(tcs_pack_reg<(V) * 4 + 0>(h0, kmask), tcs_pack_reg<(V) * 4 + 1>(h1, kmask), q1, q2, q3); \
                tcs_key_insert2_keys(tcs_pack_reg<(V) * 4 + 2>(h2, kmask), tcs_pack_reg<(V) * 4 + 3>(h3, kmask), q1, q2, q3); \
              }
              TC_SCORE4(0) TC_SCORE4(1) TC_SCORE4(2) TC_SCORE4(3)
#undef TC_SCORE4
            } else {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                tcs_key_insert(fmaf(__uint_as_float(s[v * 4 + 0]), ninv, t[v].x), v * 4 + 0, q1, q2, q3);
                tcs_key_insert(fmaf(__uint_as_float(s[v * 4 + 1]), ninv, t[v].y), v * 4 + 1, q1, q2, q3);
                tcs_key_insert(fmaf(__uint_as_float(s[v * 4 + 2]), ninv, t[v].z), v * 4 + 2, q1, q2, q3);
                tcs_key_insert(fmaf(__uint_as_float(s[v * 4 + 3]), ninv, t[v].w), v * 4 + 3, q1, q2, q3);
              }
            }
            tcs_merge(q1, q2, q3, col, m1, m2, m3, i1, i2);
          }
        };
        fold_t(ta0, tb0, col0);
        // software pipeline, two chunks per trip: the next chunk's Gram rows and TMEM columns are in flight while this one
        // is scored.  The body is ~400 instructions on purpose (see the instruction-cache note in the file header).
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          load_t(ta1, tb1, col0 + c + 16);
          tc_ld_wait();                                   // s0 landed
          tc_ld16_issue(tcol + c + 16, s1);
          score16(s0, ta0, col0 + c);
          fold_t(ta1, tb1, col0 + c + 16);
          // No `if (last trip)` around the prefetches: a branch here splits the body into basic blocks and ptxas then hoists
          // the next score16 above the prefetch block (seen in SASS), which exposes the full L2 latency again.  The last trip
          // harmlessly re-fetches the chunk it just scored.
          const int cn = min(c + 32, 96);
          load_t(ta0, tb0, col0 + cn);
          tc_ld_wait();                                   // s1 landed
          tc_ld16_issue(tcol + cn, s0);
          score16(s1, ta1, col0 + c + 16);
          fold_t(ta0, tb0, col0 + cn);
        }
        tc_ld_wait();

        TC_ACC(e_scan, te);
        if (quarter == 0) TC_EV(ev_role, 2, it * 16 + l);
        if (half == 1) {
          // ---- hand the top-3 of columns [128,256) to the half-0 warp of this lane quarter, then wait for the final id
          TcExch e; e.m1 = m1; e.m2 = m2; e.m3 = m3; e.idx = (uint32_t)i1 | ((uint32_t)i2 << 8);
          ms->exch[r_local] = e;
          release_tmem(buf);
          tc_pair_arrive(bar_x);
          tc_pair_sync(bar_i);
          TC_ACC(e_idw, te);
          if (quarter == 0) TC_EV(ev_role, 3, it * 16 + l);
          idpack |= (uint64_t)(ms->exch[r_local].idx & 0xff) << (8 * l);
          continue;
        }

        tc_pair_sync(bar_x);
        TC_ACC(e_pair, te);
        {
          const TcExch e = ms->exch[r_local];
          tc_insert(e.m1, (int)(e.idx & 0xff), m1, m2, m3, i1, i2);
          tc_insert(e.m2, (int)((e.idx >> 8) & 0xff), m1, m2, m3, i1, i2);
          m3 = fminf(m3, fmaxf(m2, e.m3));
        }
        const float thr = tcs_threshold(m1, margin);
        const bool flagged = valid && !(m2 > thr);          // >= 2 candidates (NaN/inf margins land here too)
        const bool many = flagged && !(m3 > thr);           // >= 3 candidates: rare, needs the full candidate mask
        const uint32_t fl = __ballot_sync(0xffffffffu, flagged);
        const uint32_t mn = __ballot_sync(0xffffffffu, many);
        // candidate bitmask of the rows with >= 3 candidates.  It lives in local memory (dynamically indexed) = L2 here,
        // so it is only touched on that rare path: every word is written inside `if (mn)` before the re-rank reads it.
        uint32_t mask[8];
        TC_ACC(e_merge, te);
        if (quarter == 0) TC_EV(ev_role, 3, it * 16 + l);
        if (mn) {
          // second pass over all 256 raw scores (warp-uniform branch): exact candidate bitmask for the `many` rows
          const uint32_t tall = TC_TMEM_BASE() + lane_addr + buf * 256;
#pragma unroll 1
          for (int w = 0; w < 8; ++w) {
            uint32_t mw = 0;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              const int col = w * 32 + hh * 16;
              tc_ld16_issue(tall + col, s0);
              load_t(ta0, tb0, col);
              fold_t(ta0, tb0, col);
              tc_ld_wait();
              uint32_t bits = 0;
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                bits |= (uint32_t)(!(fmaf(__uint_as_float(s0[v * 4 + 0]), ninv, ta0[v].x) > thr)) << (v * 4 + 0);
                bits |= (uint32_t)(!(fmaf(__uint_as_float(s0[v * 4 + 1]), ninv, ta0[v].y) > thr)) << (v * 4 + 1);
                bits |= (uint32_t)(!(fmaf(__uint_as_float(s0[v * 4 + 2]), ninv, ta0[v].z) > thr)) << (v * 4 + 2);
                bits |= (uint32_t)(!(fmaf(__uint_as_float(s0[v * 4 + 3]), ninv, ta0[v].w) > thr)) << (v * 4 + 3);
              }
              mw |= bits << (hh * 16);
            }
            mask[w] = mw;
          }
        }
        TC_ACC(e_many, te);
        release_tmem(buf);
        if (quarter == 0) TC_EV(ev_role, 4, it * 16 + l);                  // accumulator buffer may be overwritten by level l+2

        int my_id = i1;
        // ---- warp-cooperative exact re-rank of the flagged rows (same arithmetic as rq_simt.cu: sequential fp32 residual,
        // (xx + cc) - 2 dot, first index wins ties).  Lane covers elements 128 i + 4 lane .. +3 of a row (6 x LDG.128 per
        // row); the x row, the first prior code and both candidates are all in flight together.
        const float* ccl = p.cc + l * TC_K;
        const float* cl = p.cbf + (size_t)l * TC_K * D;
        uint32_t todo = fl;
        int n_cand = 0;
        auto ld_row = [&](const float* base, float4 (&v)[6]) {
#pragma unroll
          for (int i = 0; i < 6; ++i)
            v[i] = (i * 128 + lane4 < D) ? __ldg(reinterpret_cast<const float4*>(base + i * 128 + lane4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
#pragma unroll 1
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int rrow = __shfl_sync(0xffffffffu, row, src);
          const uint32_t idlo = __shfl_sync(0xffffffffu, (uint32_t)idpack, src);
          const uint32_t idhi = __shfl_sync(0xffffffffu, (uint32_t)(idpack >> 32), src);
          const uint64_t rid = ((uint64_t)idhi << 32) | idlo;
          const int ci1 = __shfl_sync(0xffffffffu, i1, src), ci2 = __shfl_sync(0xffffffffu, i2, src);
          const bool is_many = (mn >> src) & 1;
          const int ka = min(ci1, ci2), kb = max(ci1, ci2);
          float4 res[6], ev[6], va[6], vb[6];
          const float* xr = p.x + (int64_t)rrow * p.ldx;
          if (kVec) ld_row(xr, res);
          else {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (i * 128 + lane4 < D) {
                const float* q = xr + i * 128 + lane4;
                res[i].x = __ldg(q); res[i].y = __ldg(q + 1); res[i].z = __ldg(q + 2); res[i].w = __ldg(q + 3);
              }
            }
          }
          ld_row(cl + (size_t)ka * D, va);
          ld_row(cl + (size_t)kb * D, vb);
#pragma unroll 1
          for (int j = 0; j < l; ++j) {      // the j = 0 row is in flight together with x and both candidates
            ld_row(p.cbf + ((size_t)j * TC_K + (size_t)((rid >> (8 * j)) & 0xff)) * D, ev);
#pragma unroll
            for (int i = 0; i < 6; ++i) { res[i].x -= ev[i].x; res[i].y -= ev[i].y; res[i].z -= ev[i].z; res[i].w -= ev[i].w; }   // rqvae.py:130, level order
          }
          float xx = 0.f;
#pragma unroll
          for (int i = 0; i < 6; ++i) xx = tc_dot4(res[i], res[i], xx);
          xx = warp_sum(xx);
          float best = INFINITY;
          int besti = 0x7fffffff;
          if (!is_many) {
            // exactly two candidates, ascending index order
            float da = 0.f, db = 0.f;
#pragma unroll
            for (int i = 0; i < 6; ++i) { da = tc_dot4(res[i], va[i], da); db = tc_dot4(res[i], vb[i], db); }
            da = warp_sum(da);
            db = warp_sum(db);
            const float dist_a = (xx + __ldg(ccl + ka)) - 2.f * da;             // quantize.py:113-117
            const float dist_b = (xx + __ldg(ccl + kb)) - 2.f * db;
            best = dist_a; besti = ka;
            if (dist_b < best) { best = dist_b; besti = kb; }
            if (!(dist_a == dist_a)) besti = (dist_b == dist_b) ? kb : ci1;     // NaN distances: keep something valid
            n_cand += 2;
          } else {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              uint32_t mw = __shfl_sync(0xffffffffu, mask[c], src);
#pragma unroll 1
              while (mw) {
                const int k = c * 32 + __ffs(mw) - 1;
                mw &= mw - 1;
                ld_row(cl + (size_t)k * D, va);
                float dot = 0.f;
#pragma unroll
                for (int i = 0; i < 6; ++i) dot = tc_dot4(res[i], va[i], dot);
                dot = warp_sum(dot);
                const float dist = (xx + __ldg(ccl + k)) - 2.f * dot;
                if (dist < best) { best = dist; besti = k; }
                ++n_cand;
              }
            }
            if (besti > 255) besti = ci1;   // all-NaN row: keep the filter's pick
          }
          if (lane == src) my_id = besti;
        }
        if (p.stats && lane == 0 && fl) {
          atomicAdd(p.stats + 0, __popc(fl));
          atomicAdd(p.stats + 1, n_cand);
          atomicAdd(p.stats + 2, __popc(mn));
        }
        idpack |= (uint64_t)(my_id & 0xff) << (8 * l);
        ms->exch[r_local].idx = (uint32_t)my_id;     // publish the final id of this level to the half-1 warp
        tc_pair_arrive(bar_i);
        if (quarter == 0) TC_EV(ev_role, 5, it * 16 + l);
        if (valid) p.ids[(int64_t)row * L + l] = my_id;
      }
    }
    if (trace && quarter == 0 && lane == 0) {
      TC_ACC(e_rr, te);
      const int o = half ? 13 : 4;     // half 0 -> slots 4..8, half 1 -> slots 13..17
      tc_trace_add(p.stats, o + 0, e_tf); tc_trace_add(p.stats, o + 1, e_scan); tc_trace_add(p.stats, o + 2, half ? e_idw : e_pair);
      tc_trace_add(p.stats, o + 3, e_rr); tc_trace_add(p.stats, o + 4, clock64() - te_start);
      if (!half) { tc_trace_add(p.stats, 18, e_merge); tc_trace_add(p.stats, 19, e_many); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();     // neither CTA may exit (or free TMEM) while the other can still reach into it
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tc_dealloc2(TC_TMEM_BASE(), 512); else tc_dealloc(TC_TMEM_BASE(), 512);
  }
}

template <bool kTrace, bool kVec, bool kPair, int kOpt = 0>
static int tc_launch(const TcParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = rq_tc_kernel<kTrace, kVec, kPair, kOpt>;
  RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (kPair) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    RQB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    kern<<<grid, TC_THREADS, smem, st>>>(p);
  }
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

template <bool kPair>
static int tc_dispatch(const TcParams& p, int grid, size_t smem, cudaStream_t st, bool trace, bool vec_ok, int opt) {
  // opt: bit 0 = in-place TMA staging of x, bit 1 = fast scan arithmetic (both opt-in, vector-load instantiation only)
  if (vec_ok && opt == 1) return trace ? tc_launch<true, true, kPair, 1>(p, grid, smem, st) : tc_launch<false, true, kPair, 1>(p, grid, smem, st);
  if (vec_ok && opt == 2) return trace ? tc_launch<true, true, kPair, 2>(p, grid, smem, st) : tc_launch<false, true, kPair, 2>(p, grid, smem, st);
  if (vec_ok && opt == 3) return trace ? tc_launch<true, true, kPair, 3>(p, grid, smem, st) : tc_launch<false, true, kPair, 3>(p, grid, smem, st);
  if (trace) return vec_ok ? tc_launch<true, true, kPair>(p, grid, smem, st) : tc_launch<true, false, kPair>(p, grid, smem, st);
  return vec_ok ? tc_launch<false, true, kPair>(p, grid, smem, st) : tc_launch<false, false, kPair>(p, grid, smem, st);
}

extern "C" int rqb200_tokenize_tc_run(const float* x, int64_t ldx, int B, const void* state, int D, int K, int L,
                                      int64_t* ids, int* stats, void* stream) {
  if (!rqb200_tokenize_tc_supported(D, K, L)) {
    rqb_set_error("tokenize_tc: shape D=%d K=%d L=%d not supported", D, K, L);
    return RQB_ERR_UNSUPPORTED;
  }
  RQB_CHECK_ARG(B >= 0 && ldx >= D && ldx < (1 << 24), "tokenize_tc_run: bad shape");
  if (B == 0) return RQB_OK;
  RQB_CHECK_ARG(x && state && ids, "tokenize_tc_run: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const char* base = reinterpret_cast<const char*>(state);
  TcParams p{};
  p.x = x; p.ldx = ldx; p.B = B; p.D = D; p.L = L; p.nkc = D / TC_KC;
  p.ntiles = (B + TC_BM - 1) / TC_BM;
  p.hdr = reinterpret_cast<const TcHeader*>(base);
  p.cc = reinterpret_cast<const float*>(base + tc_off_cc(L));
  p.hcc = reinterpret_cast<const float*>(base + tc_off_hcc(L));
  p.gram = reinterpret_cast<const float*>(base + tc_off_gram(L));
  p.cbf = reinterpret_cast<const float*>(base + tc_off_cbf(L));
  p.blob = reinterpret_cast<const unsigned char*>(base + tc_off_blob(D, L));
  p.ids = ids; p.stats = stats; p.sx = 1.0f; p.one = 1;
  // both measured NEGATIVE on B200 (tools/tc_ab.py, same box: rotation +1 %, prefetch +0.5 % time): kept as opt-in knobs so
  // the measurement can be repeated, off by default
  static const int opt_rot = []() { const char* e = getenv("RQB200_TC_ROT"); return (e && e[0] == '1') ? 1 : 0; }();
  static const int opt_pf = []() { const char* e = getenv("RQB200_TC_PREFETCH"); return (e && e[0] == '1') ? 1 : 0; }();
  p.rot = opt_rot; p.prefetch = opt_pf;
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    RQB_CUDA(cudaGetDevice(&dev));
    RQB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t smem = (size_t)TC_MAX_KC * TC_ACHUNK_BYTES + TC_BSTAGES * TC_BSTAGE_BYTES + sizeof(TcSmemMisc);
  static const bool want_trace = []() { const char* e = getenv("RQB200_TC_TRACE"); return e && e[0] == '1'; }();
  const bool vec_ok = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const bool trace = want_trace && stats;       // tracing: caller passes >= 64 ints; 64-bit cycle accumulators start at stats[8]
  // 64-rows-per-CTA kernel (csrc/rq_tc64.cu: M = 128 CTA-pair MMAs, x staged by TMA): opt-in (one hardware run: identical ids, slower)
  // RQB200_TC_64=1: clusters of 2 (one pair); =4 / =8: clusters of 4 / 8 (two / four pairs sharing the codebook blocks by TMA multicast)
  static const int opt_64 = []() { const char* e = getenv("RQB200_TC_64"); return (e && e[0] == '1') ? 2 : (e && e[0] == '4') ? 4 : (e && e[0] == '8') ? 8 : 0; }();
  if (opt_64 && vec_ok && sm_count >= opt_64) return tc64_run(p, sm_count, trace, opt_64, st);
  // in-place TMA staging of x (kOpt bit 0 of rq_tc_kernel, single-CTA or pair): opt-in (first hardware run raced; fixed, fix unrun)
  static const int opt_tma = []() { const char* e = getenv("RQB200_TC_TMA"); return (e && e[0] == '1') ? 1 : 0; }();
  const bool tma = opt_tma && vec_ok;
  if (tma) {
    p.rot = 0;                                  // the in-place scheme walks the chunks in slot order
    int rc = tc_encode_2d(&p.tmapXh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.x, (uint64_t)D, (uint64_t)B, (uint64_t)ldx * 4, 32, TC_BM);
    if (rc) return rc;
  }
  static const int opt_fast = []() { const char* e = getenv("RQB200_TC_FASTSCAN"); return (e && e[0] == '1') ? 1 : 0; }();
  const int opt = (tma ? 1 : 0) | ((opt_fast && vec_ok) ? 2 : 0);
  // CTA-pair variant (cta_group::2): opt-in while it is being brought up
  static const int opt_pair = []() { const char* e = getenv("RQB200_TC_PAIR"); return (e && e[0] == '1') ? 1 : 0; }();
  if (opt_pair && p.ntiles >= 2 && sm_count >= 2) {
    int rc = tc_encode_blob_map(&p.tmapB, p.blob, L * 2 * p.nkc);
    if (rc) return rc;
    const int npairs = (p.ntiles + 1) / 2;
    const int nclusters = npairs < sm_count / 2 ? npairs : sm_count / 2;
    return tc_dispatch<true>(p, 2 * nclusters, smem, st, trace, vec_ok, opt);
  }
  const int grid = p.ntiles < sm_count ? p.ntiles : sm_count;
  return tc_dispatch<false>(p, grid, smem, st, trace, vec_ok, opt);
}
