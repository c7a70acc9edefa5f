// Candidate selection core of the tensor-core tokeniser (csrc/rq_tc.cu), written so the SAME code compiles for the
// device (nvcc) and for the host (g++: tests/test_tc_select.py drives it against a brute-force model on the CPU).
//
// Per row and level the epilogue sees 256 approximate half-distances a[k] = T[k] - S[k] * inv (quantize.py:113-117 up
// to the row constant xx/2).  It has to return the argmin candidate set {k : a[k] <= min + margin}: one candidate = the
// exact answer, two = (i1, i2) for the exact re-rank, three or more = `many` (full candidate mask, rare).
//
// Two-stage top-3:
//   stage 1 (per 16-column chunk, 7 instructions per score): the column index e (4 bits) is packed into the low
//           mantissa bits of the score, so a running sorted triple of KEYS needs only min/max, no compares or selects;
//   stage 2 (per chunk): the triple is merged into the running top-3 values / top-2 indices with tc_insert.
// The packing perturbs a score by < 16 ulp <= 2^-19 |a|; tcs_threshold() widens the margin by that amount, so the
// candidate set can only grow (more re-rank work, never a missed candidate).
#ifndef RQB200_TC_SELECT_CUH
#define RQB200_TC_SELECT_CUH
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define TCS_HD __host__ __device__ __forceinline__
#else
#define TCS_HD inline
#endif

#define TCS_CHUNK 16
#define TCS_IDX_MASK 0xFu
#define TCS_KEY_MASK 0xFFFFFFF0u

TCS_HD uint32_t tcs_f2u(float v) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(v);
#else
  uint32_t u; memcpy(&u, &v, 4); return u;
#endif
}
TCS_HD float tcs_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float v; memcpy(&v, &u, 4); return v;
#endif
}

// branch-free insertion of (a, k) into the sorted top-3 values / top-2 indices; strict '<' keeps the earlier index
TCS_HD void tc_insert(float a, int k, float& m1, float& m2, float& m3, int& i1, int& i2) {
  const bool lt1 = a < m1, lt2 = a < m2;
  m3 = fminf(m3, fmaxf(m2, a));
  m2 = fminf(m2, fmaxf(m1, a));
  m1 = fminf(m1, a);
  i2 = lt1 ? i1 : (lt2 ? k : i2);
  i1 = lt1 ? k : i1;
}

// stage 1: fold score `a` of chunk column e (0..15, compile-time after unrolling) into the sorted key triple
TCS_HD void tcs_key_insert(float a, uint32_t e, float& q1, float& q2, float& q3) {
  const float key = tcs_u2f((tcs_f2u(a) & TCS_KEY_MASK) | e);
  q3 = fminf(q3, fmaxf(q2, key));
  q2 = fminf(q2, fmaxf(q1, key));
  q1 = fminf(q1, key);
}

// stage 1, two scores per call: 8 min/max instructions instead of 10 (sm_100 has three-input FMNMX3, which ptxas forms from
// the nested fminf).  With lo <= hi the two keys and q1 <= q2 <= q3 the running triple, the k-th smallest of the merged
// lists is min over i + j = k of max(q_i, key_j):  r1 = min(q1, lo),  r2 = min(max(q1, lo), q2, hi),
// r3 = min(q3, max(q2, lo), max(q1, hi)).
TCS_HD void tcs_key_insert2_keys(float ka, float kb, float& q1, float& q2, float& q3) {   // ka, kb: packed keys
  const float lo = fminf(ka, kb), hi = fmaxf(ka, kb);
  q3 = fminf(fminf(q3, fmaxf(q2, lo)), fmaxf(q1, hi));
  q2 = fminf(fminf(fmaxf(q1, lo), q2), hi);
  q1 = fminf(q1, lo);
}
TCS_HD void tcs_key_insert2(float a, uint32_t ea, float b, uint32_t eb, float& q1, float& q2, float& q3) {
  tcs_key_insert2_keys(tcs_u2f((tcs_f2u(a) & TCS_KEY_MASK) | ea), tcs_u2f((tcs_f2u(b) & TCS_KEY_MASK) | eb), q1, q2, q3);
}
#ifdef __CUDACC__
// (a & mask) | E in ONE LOP3: with both the mask and the column index as immediates ptxas needs two instructions per key; the
// mask therefore comes in a register (the caller keeps it opaque) and only E is an immediate.  Same value as the generic form.
template <uint32_t E>
__device__ __forceinline__ float tcs_pack_reg(float a, uint32_t mask_in_register) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(mask_in_register), "n"(E));
  return __uint_as_float(r);
}
#endif

// stage 2: merge the key triple of the chunk whose first column is `cbase` into the running top-3
TCS_HD void tcs_merge(float q1, float q2, float q3, int cbase, float& m1, float& m2, float& m3, int& i1, int& i2) {
  tc_insert(q1, cbase + (int)(tcs_f2u(q1) & TCS_IDX_MASK), m1, m2, m3, i1, i2);
  tc_insert(q2, cbase + (int)(tcs_f2u(q2) & TCS_IDX_MASK), m1, m2, m3, i1, i2);
  m3 = fminf(m3, fmaxf(m2, q3));   // q3 >= q2: it can only displace m3
}

// candidate threshold on keys: every k with a[k] <= a_min + margin has key[k] <= tcs_threshold(m1, margin), m1 = min key.
// |key - a| < 16 ulp(a) <= 2^-19 |a| (normal a) on both the minimum and the candidate gives
//   key[k] <= m1 + margin (1 + 2^-19) + 2^-18 |m1| (1 + tiny);  2^-18 = 3.815e-6, the rest of 4.5e-6 covers the rounding of
// this expression itself; 1e-43 covers 16 denormal ulps when the scores themselves are denormal.
TCS_HD float tcs_threshold(float m1, float margin) {
  return m1 + (margin * 1.00001f + 4.5e-6f * fabsf(m1) + 1e-43f);
}

#endif  // RQB200_TC_SELECT_CUH
