// Tensor-core tokeniser, 64 rows per CTA: M = 128 CTA-pair MMAs, x staged through shared memory by TMA.
//
// Same result contract, same prepared state and same filter / re-rank arithmetic as rq_tc_kernel (csrc/rq_tc.cu, whose header
// states the algorithm; reference: modules/quantize.py:113-128,159-161 x L + modules/rqvae.py:125-132).  What changes is
// where the bytes live, because the 128-row kernel is out of shared memory (A resident = 192 KB of 227 KB):
//   * its x path (LDG -> registers -> fp16 -> smem) can keep only ~32 KB in flight per SM and shares the L1TEX/LSU path
//     with the epilogue's gathers: measured 5-6 B/clk/SM against an HBM share of 23 B/clk/SM, pacing level 0 of every tile;
//   * its codebook ring is 2 x 16 KB, so levels 1-2 run at the L2 latency, not at the tensor rate.
// Here a cluster of two CTAs owns a 128-row pair-tile, 64 rows each:
//   A (fp16 image of x)   12 x 8 KB = 96 KB resident per CTA (K-major SWIZZLE_128B, 64 rows)
//   x staging             TC64_NX x 16 KB: 64 rows x 64 fp32 boxes of a 2-D tensor map over x (TMA, no LSU, zero-filled past B)
//   B ring                TC64_NB x 16 KB: this CTA's 128 codes of a (level, k chunk) block (tensor-map TMA, cta_group::2)
//   MMA                   tcgen05.mma.cta_group::2.kind::f16, M = 128 (64 rows from each CTA) x N = 256 x K = 16, issued by the
//                         leader; the guide's pacing law gives it the full per-SM rate (max(M,128) N / (256 x 2) = 64 cycles)
//   accumulators          "2x2" TMEM layout (cute tmem_frg_2sm, M_MMA_SM = 64): row m of this CTA sits in lane m for codes
//                         [0,128) and in lane 64 + m for codes [128,256), so a level needs 128 columns -> FOUR accumulator
//                         buffers (the 128-row kernel has two), and every one of the 128 lanes is used
//   epilogue              8 warps: warp (quarter, sub) scans TMEM lanes [32 quarter, +32), columns [64 sub, +64) = rows
//                         32 (quarter & 1) + lane, codes 128 (quarter >> 1) + 64 sub + [0,64); the four partial top-3 of a row
//                         meet in shared memory, the (quarter < 2, sub 0) warp owns merge / re-rank / ids
// STATUS: written without GPU access at the end of round 1.  Its first (and so far only) hardware run -- commit 473cb52, clusters of
// 2 / 4 / 8, 1 000 and 65 536 rows -- returned ids byte-identical to rq_tc_kernel's and was 16 % slower; the
// TMEM layout and the full-rate M = 128 pair instruction were confirmed by tools/pair64_probe.cu in the same call.  Ring
// parameters, the named-barrier scheme, the scan arithmetic and the kGrp = 2 variant were changed AFTER that run: opt-in
// (RQB200_TC_64), gated test (RQB200_TEST_UNVALIDATED=1), index arithmetic unit-tested on the host (tests/test_tc64_layout.py).
#include "tc_common.cuh"
#include "tc64_layout.cuh"

#define TC64_BM 64                                  // rows per CTA
#define TC64_ACHUNK_BYTES (TC64_BM * TC_KC * 2)     // 8 KB
#define TC64_XSTAGE_BYTES (TC64_BM * TC_KC * 4)     // 16 KB: 64 rows x 64 fp32
#define TC64_NB 4                                   // default depth of the codebook ring   (run time: RQB200_TC64_NB, p.nb)
#define TC64_NX 3                                   // default depth of the x staging ring  (run time: RQB200_TC64_NX, p.nx)
#define TC64_MAXS 6                                 // barrier slots per ring; nb + nx <= 7 stages of 16 KB fit beside A
#define TC64_NBUF 4                                 // accumulator buffers of 128 TMEM columns

struct Tc64Misc {
  uint64_t a_full[TC_MAX_KC], a_empty[TC_MAX_KC];
  uint64_t b_full[TC64_MAXS], b_empty[TC64_MAXS];
  uint64_t b_peer[TC64_MAXS];      // kCl = 4 only, pair leader: the peer CTA's codebook stage has landed (forwarded)
  uint64_t x_full[TC64_MAXS], x_empty[TC64_MAXS];
  uint64_t t_full[TC64_NBUF], t_empty[TC64_NBUF];
  // everything below exists once per epilogue GROUP (kGrp = 2: two groups of four warps work on alternate tiles)
  uint64_t rowinfo_free[2];
  uint32_t tmem_base;
  uint32_t many_word[2][2];       // per row group: ballot of the rows with >= 3 candidates at the level being merged
  uint32_t pad;
  uint32_t rowinfo[2][TC64_BM];   // bf16x2 (rounded up): max|x| | sum x^2 of the tile being scored
  float thr[2][TC64_BM];          // candidate threshold of the level being merged (owner -> partner warps, `many` rows)
  uint32_t idpub[2][TC64_BM];     // final id of the level (owner -> partner warps)
  TcExch exch[2][3][TC64_BM];     // partner slot -> row: top-3 of that warp's columns
  uint32_t mask[2][TC64_BM][8];   // candidate bitmask of the `many` rows (each warp writes the words of its own columns)
};

// tc_tma2d (tc_common.cuh) delivered to the same shared-memory offset (data and mbarrier) of every CTA whose bit is set in `mask`
__device__ __forceinline__ void tc64_tma2d_mc(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// completion of all prior tcgen05 ops of this thread -> the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc64_commit_mask(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// position in a ring of n stages: stage index + phase parity, advanced without divisions (n is a run-time parameter)
struct Tc64Ring {
  uint32_t st, ph, n;
  __device__ __forceinline__ explicit Tc64Ring(int n_) : st(0), ph(0), n((uint32_t)n_) {}
  __device__ __forceinline__ void next() { if (++st == n) { st = 0; ph ^= 1u; } }
};
// named barriers of the warps that share the rows of one row group (4 warps, or 2 when kGrp = 2)
__device__ __forceinline__ void tc64_grp_sync(int id, int nthr) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory"); }
__device__ __forceinline__ void tc64_grp_arrive(int id, int nthr) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthr) : "memory");
}

// kCl = cluster size.  2: one CTA pair per cluster, every pair streams the codebook blocks from L2 by itself.
// 4 / 8: two / four pairs per cluster share every block: each of the CTAs that need a given 16 KB half-block loads one slice
// of it and multicasts it to all of them (1/2, 1/4 of the L2 reads per row; the pairs couple only through the B ring's depth).
// kGrp = epilogue groups.  1: all 8 epilogue warps work on the same tile (4 warps per row, 64 columns each).  2: two groups of
// 4 warps take ALTERNATE tiles (2 warps per row, 128 columns each): the dependency chain of one tile -- scan, merge, serial
// exact re-ranks, ids, next level's Gram pointers (it, not any chip-wide throughput, is what bounds the
// kernel) -- gets longer, but two chains overlap; the 4 accumulator buffers already hold both tiles' levels.
template <bool kTrace, int kCl, int kGrp>
__global__ void __launch_bounds__(TC_THREADS, 1) rq_tc64_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char tsm[];
  unsigned char* sA = tsm;                                             // [TC_MAX_KC][8 KB]
  unsigned char* sB = sA + TC_MAX_KC * TC64_ACHUNK_BYTES;              // [p.nb][16 KB]
  unsigned char* sX = sB + p.nb * TC_BSTAGE_BYTES;                     // [p.nx][16 KB]
  Tc64Misc* ms = reinterpret_cast<Tc64Misc*>(sX + p.nx * TC64_XSTAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkc = p.nkc, L = p.L;
  const bool trace = kTrace && p.stats != nullptr;
  const uint32_t rank = cluster_ctarank();
  const uint32_t crank = rank & 1u;                                   // rank inside the CTA pair, 0 = its leader
  const uint32_t leader = rank & ~1u;                                 // cluster rank of this pair's leader
  const uint32_t pair = rank >> 1;                                    // which pair of the cluster (0 when kCl == 2)
  const uint16_t pair_mask = (uint16_t)(3u << leader);                // both CTAs of this pair
  // work unit = kCl consecutive 64-row tiles (one cluster); this CTA takes tile kCl * unit + rank
  const int u_first = (int)(blockIdx.x / kCl), u_step = (int)(gridDim.x / kCl);
  const int ntiles64 = (p.B + TC64_BM - 1) / TC64_BM;
  const int u_count = (ntiles64 + kCl - 1) / kCl;

  if (tid == 0) {
    if ((smem_u32(tsm) & 1023u) != 0) __trap();  // the swizzle pattern needs a 1024-byte aligned base
    for (int i = 0; i < TC_MAX_KC; ++i) { mbar_init(&ms->a_full[i], 2 * TC_NCONV_WARPS); mbar_init(&ms->a_empty[i], 1); }
    for (int i = 0; i < TC64_MAXS; ++i) {
      mbar_init(&ms->b_full[i], 1);
      mbar_init(&ms->b_empty[i], kCl / 2);   // one multicast commit per pair leader that reads the (shared) stage
      mbar_init(&ms->b_peer[i], 1);
    }
    for (int i = 0; i < TC64_MAXS; ++i) { mbar_init(&ms->x_full[i], 1); mbar_init(&ms->x_empty[i], TC_NCONV_WARPS); }
    for (int i = 0; i < TC64_NBUF; ++i) { mbar_init(&ms->t_full[i], 1); mbar_init(&ms->t_empty[i], 2 * TC_NEPI_WARPS / kGrp); }
    mbar_init(&ms->rowinfo_free[0], TC64_BM);    // the owner thread of every row
    mbar_init(&ms->rowinfo_free[1], TC64_BM);
    fence_mbar_init();
  }
  if (warp == 1) tc_alloc2(&ms->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // the peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  #define TC64_TMEM_BASE() (*reinterpret_cast<volatile uint32_t*>(&ms->tmem_base))

  if (warp < 4) {
    // ============================================================== warpgroup 0: B producer, MMA issuer, x producer
    tc_setmaxnreg_dec<32>();
    if (warp == 0) {
      // this CTA's 128 codes (column half = crank) of every (level, k chunk) block; the bytes of both CTAs are counted on
      // the LEADER's b_full, which is what its MMA warp waits on
      Tc64Ring rb(p.nb);
      for (int unit = u_first; unit < u_count; unit += u_step)
        for (int l = 0; l < L; ++l)
          for (int kc = 0; kc < nkc; ++kc, rb.next()) {
            const uint32_t st = rb.st;
            mbar_wait_guarded(&ms->b_empty[st], rb.ph ^ 1u, 1);       // local: the leaders' commits are multicast
            if (tc_elect_one()) {
              const int blk_row = ((l * 2 + (int)crank) * nkc + kc) * 128;      // first blob row of this CTA's half-block
              if constexpr (kCl == 2) {
                if (crank == 0) mbar_expect_tx(&ms->b_full[st], 2 * TC_BSTAGE_BYTES);
                tc_tma2d_pair(sB + st * TC_BSTAGE_BYTES, &p.tmapB, 0, blk_row, cluster_map(smem_u32(&ms->b_full[st]), leader));
              } else {
                // this CTA and the CTAs of the other pairs with the same pair-rank need the same half-block: each loads
                // slice `pair` of it and multicasts it to all of them; every CTA arms its OWN b_full for the 16 KB it receives
                // (a slice may land before the receiver has armed this phase: the pending arrival keeps the phase open)
                constexpr int kSliceRows = 128 / (kCl / 2);             // 64 (8 KB) for clusters of 4, 32 (4 KB) for clusters of 8
                mbar_expect_tx(&ms->b_full[st], TC_BSTAGE_BYTES);
                tc64_tma2d_mc(sB + st * TC_BSTAGE_BYTES + pair * (kSliceRows * 128), &p.tmapB2, 0, blk_row + (int)pair * kSliceRows,
                              &ms->b_full[st], (uint16_t)((0x55u & ((1u << kCl) - 1u)) << crank));
              }
            }
            __syncwarp();
          }
    } else if (warp == 1 && crank == 0) {
      const uint32_t idesc = tc_idesc(128, 256);
      const uint16_t all_mask = (uint16_t)((1u << kCl) - 1u);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      uint32_t g = 0, it = 0;
      Tc64Ring rb(p.nb);
      long long w_te = 0, w_af = 0, w_bf = 0, w_issue = 0;     // RQB200_TC_TRACE=1: where the issuer's cycles go
      TC_EV_DECL();
      TC_T0(tm);
      const long long tm_start = tm;
      for (int unit = u_first; unit < u_count; unit += u_step, ++it)
        for (int l = 0; l < L; ++l, ++g) {
          const uint32_t buf = g % TC64_NBUF, u = g / TC64_NBUF;
          TC_ACC(w_issue, tm);
          mbar_wait_guarded_cluster(&ms->t_empty[buf], (u & 1) ^ 1, 2);
          TC_ACC(w_te, tm);
          TC_EV(0, 1, it * 16 + l);
          tc_fence_after();
          const uint32_t d_tmem = TC64_TMEM_BASE() + buf * 128;
          for (int kc = 0; kc < nkc; ++kc, rb.next()) {
            TC_ACC(w_issue, tm);
            if (l == 0) {
              mbar_wait_guarded_cluster(&ms->a_full[kc], it & 1, 3);
              TC_EV(0, 2, it * 16 + kc);
            }
            TC_ACC(w_af, tm);
            const uint32_t st = rb.st;
            mbar_wait_guarded_cluster(&ms->b_full[st], rb.ph, 4);
            if constexpr (kCl > 2) mbar_wait_guarded_cluster(&ms->b_peer[st], rb.ph, 10);
            TC_ACC(w_bf, tm);
            tc_fence_after();
            const uint64_t adesc = tc_smem_desc(a_base + kc * TC64_ACHUNK_BYTES);
            const uint64_t bdesc = tc_smem_desc(b_base + st * TC_BSTAGE_BYTES);
            if (tc_elect_one()) {
#pragma unroll
              for (int j = 0; j < TC_KC / 16; ++j)   // K=16 per instruction: +32 B inside the 128 B swizzle row
                tc_mma_f16_2(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (kc | j) != 0);
              tc64_commit_mask(&ms->b_empty[st], all_mask);       // every CTA whose producer writes into a stage we read
              if (l == L - 1) tc64_commit_mask(&ms->a_empty[kc], pair_mask);
              if (kc == nkc - 1) tc64_commit_mask(&ms->t_full[buf], pair_mask);
            }
            __syncwarp();
          }
          TC_EV(0, 3, it * 16 + l);
        }
      if (trace && lane == 0) {      // same slots as rq_tc_kernel (tools/trace_tc.py, tools/tc_native_check.cu)
        tc_trace_add(p.stats, 0, w_te); tc_trace_add(p.stats, 1, w_af); tc_trace_add(p.stats, 2, w_bf);
        tc_trace_add(p.stats, 3, clock64() - tm_start); tc_trace_add(p.stats, 12, 1);
      }
    } else if (warp == 2) {
      // x producer: one 64-row x 64-float box per k chunk.  Rows past B read as zero (tensor-map bounds); a pair's second
      // CTA past the last 64-row tile loads the last tile again (its scores are never stored).
      Tc64Ring rx(p.nx);
      for (int unit = u_first; unit < u_count; unit += u_step) {
        const int tile = min(kCl * unit + (int)rank, ntiles64 - 1);
        for (int kc = 0; kc < nkc; ++kc, rx.next()) {
          const uint32_t st = rx.st;
          mbar_wait_guarded(&ms->x_empty[st], rx.ph ^ 1u, 8);
          if (tc_elect_one()) {
            mbar_expect_tx(&ms->x_full[st], TC64_XSTAGE_BYTES);
            tc_tma2d(sX + st * TC64_XSTAGE_BYTES, &p.tmapX, kc * TC_KC, tile * TC64_BM, &ms->x_full[st]);
          }
          __syncwarp();
        }
      }
    } else if (kCl > 2 && warp == 3 && crank == 1) {
      // the leader's MMA warp reads the codebook stage of BOTH CTAs of the pair; with multicast loads each CTA's bytes are
      // counted on its own b_full, so the peer forwards every completed phase to the leader
      Tc64Ring rb(p.nb);
      for (int unit = u_first; unit < u_count; unit += u_step)
        for (int i = 0; i < L * nkc; ++i, rb.next()) {
          const uint32_t st = rb.st;
          mbar_wait_guarded(&ms->b_full[st], rb.ph, 11);
          if (lane == 0) mbar_arrive_cluster(cluster_map(smem_u32(&ms->b_peer[st]), leader));
          __syncwarp();
        }
    }
  } else if (warp < 4 + TC_NCONV_WARPS) {
    // ============================================================== warpgroup 1: fp32 staging -> fp16 swizzled A chunks
    // Warp cw owns rows [16cw, 16cw+16) of every chunk; step j covers rows 16cw + 2j + (lane >> 4), float4 column lane & 15:
    // a warp reads 512 contiguous staging bytes (LDS.128, conflict-free) and writes two 128-byte A rows (STS.64).
    const int cw = warp - 4;
    const int hi = lane >> 4, q = lane & 15;
    uint32_t it = 0;
    Tc64Ring rx(p.nx);
    long long c_wx = 0, c_wa = 0;     // RQB200_TC_TRACE=1: waiting for the x stage / for the A slot
    TC_EV_DECL();
    const long long tcv_start = trace ? clock64() : 0;
#pragma unroll 1
    for (int unit = u_first; unit < u_count; unit += u_step, ++it) {
      float sm[8], s2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { sm[j] = 0.f; s2[j] = 0.f; }
#pragma unroll 1
      for (int kc = 0; kc < nkc; ++kc, rx.next()) {
        const uint32_t st = rx.st;
        {
          const long long t_in = trace ? clock64() : 0;
          mbar_wait_guarded(&ms->x_full[st], rx.ph, 9);
          const long long t_x = trace ? clock64() : 0;
          mbar_wait_guarded(&ms->a_empty[kc], (it & 1) ^ 1, 5);   // the last level of the previous tile released this chunk
          if (trace) { c_wx += t_x - t_in; c_wa += clock64() - t_x; }
        }
        if (cw == 0) TC_EV(1, 1, it * 16 + kc);
        const unsigned char* xs = sX + st * TC64_XSTAGE_BYTES;
        const uint32_t a_chunk = smem_u32(sA) + kc * TC64_ACHUNK_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 16 * cw + 2 * j + hi;
          const float4 a = *reinterpret_cast<const float4*>(xs + tc64_stage_offset(r, q));
          s2[j] = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, s2[j]))));
          sm[j] = fmaxf(fmaxf(sm[j], fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
          const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_chunk + tc64_a_offset(r, q)),
                       "r"(*reinterpret_cast<const uint32_t*>(&h0)), "r"(*reinterpret_cast<const uint32_t*>(&h1)) : "memory");
        }
        if (kc == nkc - 1) {
          // row statistics for the margin: reduce over the 16 lanes that share a row, publish before the last arrive
          const uint32_t gi = kGrp == 2 ? (it & 1u) : 0u;                    // which group will score this tile
          const uint32_t gu = kGrp == 2 ? (it >> 1) : it;                    // how many tiles that group has taken before
          mbar_wait_guarded(&ms->rowinfo_free[gi], (gu & 1u) ^ 1u, 6);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
              sm[j] = fmaxf(sm[j], __shfl_xor_sync(0xffffffffu, sm[j], o));
              s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
            if (q == 0) ms->rowinfo[gi][16 * cw + 2 * j + hi] = (tc_bf16_up(sm[j]) << 16) | tc_bf16_up(s2[j]);
          }
        }
        fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ms->x_empty[st]);                                       // staging stage may be refilled
          mbar_arrive_cluster(cluster_map(smem_u32(&ms->a_full[kc]), leader)); // pair leader: this warp's 16 rows of chunk kc are in
        }
        if (cw == 0) TC_EV(1, 2, it * 16 + kc);
      }
    }
    if (trace && cw == 0 && lane == 0) {
      tc_trace_add(p.stats, 9, c_wa); tc_trace_add(p.stats, 20, c_wx); tc_trace_add(p.stats, 10, clock64() - tcv_start);
    }
  } else {
    // ============================================================== warpgroups 2-3: scores -> candidates -> exact re-rank -> ids
    tc_setmaxnreg_inc<176>();
    constexpr int NSUB = 2 / kGrp;                      // column ranges per 128-column lane half
    constexpr int COLS = 128 / NSUB;                    // columns a thread scans: 64 (kGrp = 1) or 128 (kGrp = 2)
    constexpr int NP = 2 * NSUB - 1;                    // partner warps per row: 3 or 1
    constexpr int NTHR = 32 * (NP + 1);                 // threads of the named barriers below
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
    const int whi = (warp - (4 + TC_NCONV_WARPS)) >> 2; // upper / lower four epilogue warps
    const int sub = kGrp == 1 ? whi : 0;                // which COLS of the 128 columns
    const int grp = kGrp == 1 ? 0 : whi;                // which epilogue group (takes tiles it = grp, grp + kGrp, ...)
    const int rowgrp = quarter & 1, chalf = quarter >> 1;
    const int r_local = rowgrp * 32 + lane;
    const int cb = chalf * 128 + sub * COLS;            // first code this thread scores (tc64_layout.cuh for kGrp = 1)
    const bool owner = (chalf == 0 && sub == 0);
    const int slot = chalf * NSUB + sub - 1;            // partner slot 0..NP-1 (unused by the owner)
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    // three named barriers per (group, row group); every phase needs the owner AND the partners, and consecutive phases on the
    // same id are separated by a phase on another id that the other side must join first, so no phase can complete early
    const int bar_a = 2 + (grp * 2 + rowgrp) * 3;   // partners -> owner: exch[] written;  later, partners -> owner: mask[] words written
    const int bar_b = bar_a + 1;                    // owner -> partners: thr[] / many_word written
    const int bar_c = bar_a + 2;                    // owner -> partners: the level's id is final
    const int D = p.D;
    const int lane4 = lane * 4;
    TC_EV_DECL();
    const int ev_role = 2 + whi;        // lane quarter 0 only
    long long e_tf = 0, e_scan = 0, e_part = 0, e_rr = 0;   // RQB200_TC_TRACE=1 (owner warps of lane quarter 0)
    TC_T0(te);
    const long long te_start = te;
    auto release_tmem = [&](uint32_t buf) {   // one arrive per warp on the LEADER's barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(cluster_map(smem_u32(&ms->t_empty[buf]), leader));
    };
#pragma unroll 1
    for (uint32_t it = (uint32_t)grp; (int)(u_first + it * u_step) < u_count; it += kGrp) {
      const int unit = u_first + (int)it * u_step;
      const int tile = kCl * unit + (int)rank;
      const int row = tile * TC64_BM + r_local;
      const bool valid = row < p.B;    // rows past B run the same code on zero scores; nothing of theirs is stored
      uint64_t idpack = 0;             // 8 bits per level
      float x4s = 0.f, x2s = 0.f;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const uint32_t g = it * (uint32_t)L + (uint32_t)l;      // position in the MMA issuer's level sequence
        const uint32_t buf = g % TC64_NBUF, u = g / TC64_NBUF;
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
        auto fold_t = [&](float4 (&ta)[4], const float4 (&tb)[4], int col) {
          if (l >= 2) {
#pragma unroll
            for (int v = 0; v < 4; ++v) { tc_add2(ta[v].x, ta[v].y, tb[v].x, tb[v].y); tc_add2(ta[v].z, ta[v].w, tb[v].z, tb[v].w); }
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
        float4 ta0[4], tb0[4], ta1[4], tb1[4];
        load_t(ta0, tb0, cb);            // in flight across the accumulator wait below
        TC_ACC(e_rr, te);
        mbar_wait_guarded(&ms->t_full[buf], u & 1, 7);
        TC_ACC(e_tf, te);
        if (quarter == 0) TC_EV(ev_role, 1, it * 16 + l);
        tc_fence_after();
        const uint32_t tcol = TC64_TMEM_BASE() + lane_addr + buf * 128 + sub * COLS;
        uint32_t s0[16], s1[16];
        tc_ld16_issue(tcol, s0);
        const TcLevelConst lc = p.hdr->lv[l];
        if (l == 0 && owner) {           // only the merging warp needs the margin
          const uint32_t ri = ms->rowinfo[grp][r_local];
          const float xmax = __uint_as_float(ri & 0xffff0000u);        // max|x| (bf16, rounded up)
          x2s = __uint_as_float(ri << 16);                              // sum x^2 (bf16, rounded up; NaN if any input is)
          x4s = (xmax * p.sx < 65504.f) ? xmax * xmax * x2s : INFINITY;  // sum x^4 <= max|x|^2 sum x^2; fp16 overflow/inf -> poison
          mbar_arrive(&ms->rowinfo_free[grp]);
        }
        // ---- margin (DESIGN.md "filter error bound"), identical to rq_tc_kernel
        const float x2n = sqrtf(x2s);
        const float sig = 4.8828125e-4f * 0.81649658f * sqrtf(sqrtf(x4s) * lc.c4max);          // u=2^-11, sqrt(2/3)
        const float flo = 2.98023224e-8f * (lc.c1max / p.sx + sqrtf((float)p.D) * x2n / lc.sc);  // fp16 subnormal floor
        const float acc = 7.62939453e-6f * x2n * lc.c2max;                                      // 64 * 2^-23 accumulate
        const float eps = TC_Z * sig + flo + acc + lc.gerr;
        const float margin = 2.f * eps;
        const float ninv = -1.f / (p.sx * lc.sc);

        float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
        int i1 = 0, i2 = 0;
        const uint32_t kmask = p.one ? TCS_KEY_MASK : 0u;      // the key mask in a REGISTER (tcs_pack_reg)
        auto score16 = [&](const uint32_t (&s)[16], const float4 (&t)[4], int col) {
          if (p.one) {   // opaque always-true branch = basic-block boundary (see rq_tc.cu: keeps the prefetches early in SASS)
            float q1 = INFINITY, q2 = INFINITY, q3 = INFINITY;
            // per float4 of Gram/T values: two FFMA2 (two scores each), four one-instruction key packs, two pair insertions
#define TC64_SCORE4(V)                                                                                                   \
            {                                                                                                              \
              float h0, h1, h2, h3;                                                                                        \
              tc_fma2(h0, h1, __uint_as_float(s[(V) * 4 + 0]), __uint_as_float(s[(V) * 4 + 1]), ninv, t[V].x, t[V].y);   \
              tc_fma2(h2, h3, __uint_as_float(s[(V) * 4 + 2]), __uint_as_float(s[(V) * 4 + 3]), ninv, t[V].z, t[V].w);   \
              tcs_key_insert2_keys(tcs_pack_reg<(V) * 4 + 0>(h0, kmask), tcs_pack_reg<(V) * 4 + 1>(h1, kmask), q1, q2, q3); \
              tcs_key_insert2_keys(tcs_pack_reg<(V) * 4 + 2>(h2, kmask), tcs_pack_reg<(V) * 4 + 3>(h3, kmask), q1, q2, q3); \
            }
            TC64_SCORE4(0) TC64_SCORE4(1) TC64_SCORE4(2) TC64_SCORE4(3)
#undef TC64_SCORE4
            tcs_merge(q1, q2, q3, col, m1, m2, m3, i1, i2);
          }
        };
        fold_t(ta0, tb0, cb);
        // software pipeline over the thread's COLS columns, two 16-column chunks per trip (same body as rq_tc_kernel)
#pragma unroll 1
        for (int c = 0; c < COLS; c += 32) {
          load_t(ta1, tb1, cb + c + 16);
          tc_ld_wait();                                   // s0 landed
          tc_ld16_issue(tcol + c + 16, s1);
          score16(s0, ta0, cb + c);
          fold_t(ta1, tb1, cb + c + 16);
          const int cn = min(c + 32, COLS - 32);          // the last trip harmlessly re-fetches its own first chunk (no branch: see rq_tc.cu)
          load_t(ta0, tb0, cb + cn);
          tc_ld_wait();                                   // s1 landed
          tc_ld16_issue(tcol + cn, s0);
          score16(s1, ta1, cb + c + 16);
          fold_t(ta0, tb0, cb + cn);
        }
        tc_ld_wait();
        TC_ACC(e_scan, te);
        if (quarter == 0) TC_EV(ev_role, 2, it * 16 + l);

        // exact candidate bitmask of this warp's columns for the rows of the group with >= 3 candidates (warp-uniform
        // call: tcgen05.ld is .aligned); rows that are not `many` skip the global store only
        auto many_mask = [&](float thr, bool is_many) {
          const uint32_t tall = TC64_TMEM_BASE() + lane_addr + buf * 128 + sub * COLS;
#pragma unroll 1
          for (int w = 0; w < COLS / 32; ++w) {
            uint32_t mw = 0;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              const int c = w * 32 + hh * 16;
              tc_ld16_issue(tall + c, s0);
              load_t(ta0, tb0, cb + c);
              fold_t(ta0, tb0, cb + c);
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
            if (is_many) ms->mask[grp][r_local][(cb >> 5) + w] = mw;
          }
        };

        if (!owner) {
          // ---- hand this warp's top-3 to the owner, then follow its verdict
          TcExch e; e.m1 = m1; e.m2 = m2; e.m3 = m3; e.idx = (uint32_t)i1 | ((uint32_t)i2 << 8);
          ms->exch[grp][slot][r_local] = e;
          tc64_grp_arrive(bar_a, NTHR);
          tc64_grp_sync(bar_b, NTHR);
          const uint32_t mw = *reinterpret_cast<volatile uint32_t*>(&ms->many_word[grp][rowgrp]);
          if (mw) {
            many_mask(*reinterpret_cast<volatile float*>(&ms->thr[grp][r_local]), (mw >> lane) & 1);
            tc64_grp_arrive(bar_a, NTHR);
          }
          release_tmem(buf);
          tc64_grp_sync(bar_c, NTHR);
          if (quarter == 0) TC_EV(ev_role, 3, it * 16 + l);
          idpack |= (uint64_t)(*reinterpret_cast<volatile uint32_t*>(&ms->idpub[grp][r_local]) & 0xff) << (8 * l);
          continue;
        }

        tc64_grp_sync(bar_a, NTHR);
        TC_ACC(e_part, te);
#pragma unroll
        for (int sidx = 0; sidx < NP; ++sidx) {
          const TcExch e = ms->exch[grp][sidx][r_local];
          tc_insert(e.m1, (int)(e.idx & 0xff), m1, m2, m3, i1, i2);
          tc_insert(e.m2, (int)((e.idx >> 8) & 0xff), m1, m2, m3, i1, i2);
          m3 = fminf(m3, fmaxf(m2, e.m3));
        }
        const float thr = tcs_threshold(m1, margin);
        const bool flagged = valid && !(m2 > thr);          // >= 2 candidates (NaN/inf margins land here too)
        const bool many = flagged && !(m3 > thr);           // >= 3 candidates: rare, needs the full candidate mask
        const uint32_t fl = __ballot_sync(0xffffffffu, flagged);
        const uint32_t mn = __ballot_sync(0xffffffffu, many);
        ms->thr[grp][r_local] = thr;
        if (lane == 0) ms->many_word[grp][rowgrp] = mn;
        tc64_grp_arrive(bar_b, NTHR);
        if (quarter == 0) TC_EV(ev_role, 3, it * 16 + l);
        if (mn) {
          many_mask(thr, many);
          tc64_grp_sync(bar_a, NTHR);    // the partners' mask words are in
        }
        release_tmem(buf);
        if (quarter == 0) TC_EV(ev_role, 4, it * 16 + l);

        int my_id = i1;
        // ---- warp-cooperative exact re-rank of the flagged rows: identical arithmetic to rq_tc_kernel / rq_simt.cu
        // (sequential fp32 residual, (xx + cc) - 2 dot, first index wins ties)
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
          auto prior = [&](int j) { return p.cbf + ((size_t)j * TC_K + (size_t)((rid >> (8 * j)) & 0xff)) * D; };
          // the x row, both candidates AND the first prior code are requested before anything is consumed: one L2 round trip
          // less on this serial, critical-path step (a second prefetched prior row was tried: it spills the scan loop)
          ld_row(p.x + (int64_t)rrow * p.ldx, res);        // this kernel is only launched on 16-byte aligned rows
          ld_row(cl + (size_t)ka * D, va);
          ld_row(cl + (size_t)kb * D, vb);
          if (l >= 1) ld_row(prior(0), ev);
#pragma unroll 1
          for (int j = 0; j < l; ++j) {
            if (j >= 1) ld_row(prior(j), ev);
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
              uint32_t mw = *reinterpret_cast<volatile uint32_t*>(&ms->mask[grp][rowgrp * 32 + src][c]);
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
        ms->idpub[grp][r_local] = (uint32_t)my_id;     // publish the final id of this level to the partner warps
        tc64_grp_arrive(bar_c, NTHR);
        if (quarter == 0) TC_EV(ev_role, 5, it * 16 + l);
        if (valid) p.ids[(int64_t)row * L + l] = my_id;
      }
    }
    if (trace && quarter == 0 && owner && lane == 0) {     // one owner warp per group: slots 4..8 (group 0) / 13..17 (group 1)
      TC_ACC(e_rr, te);
      const int o = whi ? 13 : 4;
      tc_trace_add(p.stats, o + 0, e_tf); tc_trace_add(p.stats, o + 1, e_scan); tc_trace_add(p.stats, o + 2, e_part);
      tc_trace_add(p.stats, o + 3, e_rr); tc_trace_add(p.stats, o + 4, clock64() - te_start);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // neither CTA may exit (or free TMEM) while the other can still reach into it
  if (warp == 1) {
    tc_fence_after();
    tc_dealloc2(TC64_TMEM_BASE(), 512);
  }
}

template <bool kTrace, int kCl, int kGrp>
static int tc64_launch(const TcParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = rq_tc64_kernel<kTrace, kCl, kGrp>;
  RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  RQB_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

template <int kCl>
static int tc64_pick(const TcParams& p, int grid, size_t smem, cudaStream_t st, bool trace, int groups) {
  if (groups == 2) return trace ? tc64_launch<true, kCl, 2>(p, grid, smem, st) : tc64_launch<false, kCl, 2>(p, grid, smem, st);
  return trace ? tc64_launch<true, kCl, 1>(p, grid, smem, st) : tc64_launch<false, kCl, 1>(p, grid, smem, st);
}

// p: everything filled in by rqb200_tokenize_tc_run except the tensor maps.  Requires 16-byte aligned rows (x & 15 == 0,
// ldx % 4 == 0): the tensor map over x needs it, and the caller routes other inputs to rq_tc_kernel.
int tc64_run(TcParams& p, int sm_count, bool trace, int cluster, cudaStream_t st) {
  const int nblocks = p.L * 2 * p.nkc;
  int rc = tc_encode_blob_map(&p.tmapB, p.blob, nblocks);
  if (rc) return rc;
  rc = tc_encode_2d(&p.tmapB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, p.blob, 64, (uint64_t)nblocks * 128, 128, 64,
                    cluster == 8 ? 32 : 64);   // multicast slices: 8 KB (clusters of 4) or 4 KB (clusters of 8)
  if (rc) return rc;
  rc = tc_encode_2d(&p.tmapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.x, (uint64_t)p.D, (uint64_t)p.B, (uint64_t)p.ldx * 4,
                    TC_KC, TC64_BM);
  if (rc) return rc;
  const int ntiles64 = (p.B + TC64_BM - 1) / TC64_BM;
  // ring depths: defaults, or RQB200_TC64_NB / RQB200_TC64_NX for same-box sweeps (nb + nx <= 7 stages fit beside A)
  static const int env_nb = []() { const char* e = getenv("RQB200_TC64_NB"); return e ? atoi(e) : TC64_NB; }();
  static const int env_nx = []() { const char* e = getenv("RQB200_TC64_NX"); return e ? atoi(e) : TC64_NX; }();
  if (env_nb < 1 || env_nb > TC64_MAXS || env_nx < 1 || env_nx > TC64_MAXS || env_nb + env_nx > 7) {
    rqb_set_error("tokenize_tc: RQB200_TC64_NB=%d / RQB200_TC64_NX=%d out of range (1..%d each, sum <= 7)", env_nb, env_nx, TC64_MAXS);
    return RQB_ERR_INVALID;
  }
  p.nb = env_nb; p.nx = env_nx;
  // epilogue groups (kGrp): 1 = all eight epilogue warps on one tile, 2 = two groups of four on alternate tiles
  static const int groups = []() { const char* e = getenv("RQB200_TC64_GROUPS"); return (e && e[0] == '2') ? 2 : 1; }();
  const size_t smem = (size_t)TC_MAX_KC * TC64_ACHUNK_BYTES + (size_t)p.nb * TC_BSTAGE_BYTES +
                      (size_t)p.nx * TC64_XSTAGE_BYTES + sizeof(Tc64Misc);
  if ((cluster == 4 || cluster == 8) && ntiles64 > 2) {
    // how many clusters of this size (one CTA per SM at this shared-memory size) the device can hold at once: GPC boundaries
    // make this less than sm_count / cluster (queried, not assumed)
    static int max_cl[9] = {0};
    if (max_cl[cluster] == 0) {
      const void* kern = cluster == 4 ? (const void*)rq_tc64_kernel<false, 4, 1> : (const void*)rq_tc64_kernel<false, 8, 1>;
      RQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(sm_count / cluster * cluster));
      cfg.blockDim = dim3(TC_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      RQB_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      max_cl[cluster] = n > 0 ? n : 1;
    }
    const int units = (ntiles64 + cluster - 1) / cluster;
    const int ncl = units < max_cl[cluster] ? units : max_cl[cluster];
    if (cluster == 4) return tc64_pick<4>(p, 4 * ncl, smem, st, trace, groups);
    return tc64_pick<8>(p, 8 * ncl, smem, st, trace, groups);
  }
  const int units = (ntiles64 + 1) / 2;
  const int nclusters = units < sm_count / 2 ? units : sm_count / 2;
  return tc64_pick<2>(p, 2 * nclusters, smem, st, trace, groups);
}
