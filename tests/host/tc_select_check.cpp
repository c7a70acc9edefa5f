// Host-side property check of csrc/tc_select.cuh (driven by tests/test_tc_select.py).
// Brute force vs the two-stage packed-key top-3 over random 256-score rows: the candidate set must never lose a member.
#include "../../rq_vae_recommender_b200/csrc/tc_select.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>

struct Result { float m1, m2, m3; int i1, i2; };

// mode 1 mirrors rq_tc64_kernel: four warps scan 64 columns each, two scores per insertion (tcs_key_insert2), then the
// three partner triples are merged into the owner's
static Result scan_row64(const float* a) {
  Result h[4];
  for (int part = 0; part < 4; ++part) {
    float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY; int i1 = 0, i2 = 0;
    for (int c = 0; c < 64; c += TCS_CHUNK) {
      float q1 = INFINITY, q2 = INFINITY, q3 = INFINITY;
      for (int e = 0; e < TCS_CHUNK; e += 2)
        tcs_key_insert2(a[part * 64 + c + e], (uint32_t)e, a[part * 64 + c + e + 1], (uint32_t)(e + 1), q1, q2, q3);
      tcs_merge(q1, q2, q3, part * 64 + c, m1, m2, m3, i1, i2);
    }
    h[part] = Result{m1, m2, m3, i1, i2};
  }
  Result r = h[0];
  for (int part = 1; part < 4; ++part) {
    tc_insert(h[part].m1, h[part].i1, r.m1, r.m2, r.m3, r.i1, r.i2);
    tc_insert(h[part].m2, h[part].i2, r.m1, r.m2, r.m3, r.i1, r.i2);
    r.m3 = fminf(r.m3, fmaxf(r.m2, h[part].m3));
  }
  return r;
}

// mode 0 mirrors rq_tc_kernel: two warps scan 128 columns each in 16-column chunks, then the halves are merged (rq_tc.cu)
static Result scan_row(const float* a) {
  Result h[2];
  for (int half = 0; half < 2; ++half) {
    float m1 = INFINITY, m2 = INFINITY, m3 = INFINITY; int i1 = 0, i2 = 0;
    for (int c = 0; c < 128; c += TCS_CHUNK) {
      float q1 = INFINITY, q2 = INFINITY, q3 = INFINITY;
      for (int e = 0; e < TCS_CHUNK; ++e) tcs_key_insert(a[half * 128 + c + e], (uint32_t)e, q1, q2, q3);
      tcs_merge(q1, q2, q3, half * 128 + c, m1, m2, m3, i1, i2);
    }
    h[half] = Result{m1, m2, m3, i1, i2};
  }
  Result r = h[0];
  tc_insert(h[1].m1, h[1].i1, r.m1, r.m2, r.m3, r.i1, r.i2);
  tc_insert(h[1].m2, h[1].i2, r.m1, r.m2, r.m3, r.i1, r.i2);
  r.m3 = fminf(r.m3, fmaxf(r.m2, h[1].m3));
  return r;
}

int main(int argc, char** argv) {
  const int trials = argc > 1 ? atoi(argv[1]) : 200000;
  const int mode = argc > 2 ? atoi(argv[2]) : 0;
  std::mt19937 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(0.f, 1.f);
  long bad = 0, flagged_n = 0, many_n = 0, extra = 0;
  std::vector<float> a(256);
  for (int t = 0; t < trials; ++t) {
    // score families: wide spread, offset (all positive / all negative), tight clusters around the minimum, exact ties,
    // tiny magnitudes, huge magnitudes
    const int fam = t % 8;
    const float scale = (fam == 5) ? 1e-30f : (fam == 6) ? 1e30f : (fam == 7) ? 1e-41f : expf(6.f * ud(rng) - 3.f);
    const float off = (fam == 1) ? 40.f * scale : (fam == 2) ? -40.f * scale : 0.f;
    for (int k = 0; k < 256; ++k) a[k] = off + scale * nd(rng);
    const float spread = scale;
    float margin = spread * expf(-12.f * ud(rng));                 // from ~spread down to 6e-6 spread
    if (fam == 3) {                                                // cluster a few codes right at the minimum
      const int kmin = (int)(std::min_element(a.begin(), a.end()) - a.begin());
      const int n = 1 + (int)(rng() % 4);
      for (int j = 0; j < n; ++j) a[rng() % 256] = a[kmin] + margin * (2.f * ud(rng) - 0.5f);
    }
    if (fam == 4) {                                                // exact ties
      const int kmin = (int)(std::min_element(a.begin(), a.end()) - a.begin());
      a[rng() % 256] = a[kmin];
      if (rng() & 1) a[rng() % 256] = a[kmin];
      if (rng() & 1) margin = 0.f;
    }
    const Result r = mode ? scan_row64(a.data()) : scan_row(a.data());
    const float thr = tcs_threshold(r.m1, margin);
    const bool flagged = !(r.m2 > thr), many = flagged && !(r.m3 > thr);
    // brute force
    float amin = INFINITY; int kmin = 0;
    for (int k = 0; k < 256; ++k) if (a[k] < amin) { amin = a[k]; kmin = k; }
    std::vector<int> cand;
    for (int k = 0; k < 256; ++k) if (a[k] <= amin + margin) cand.push_back(k);
    bool ok = true;
    if (cand.size() == 1) {
      // unflagged rows must return exactly the argmin; flagging it anyway is allowed (extra work), losing it is not
      if (!flagged) ok = (r.i1 == kmin);
      else if (!many) ok = (r.i1 == kmin || r.i2 == kmin);
    } else if (cand.size() == 2) {
      if (!flagged) ok = false;
      else if (!many) ok = ((r.i1 == cand[0] && r.i2 == cand[1]) || (r.i1 == cand[1] && r.i2 == cand[0]));
    } else {
      if (!many) ok = false;
    }
    // the many path rebuilds the set from raw scores with the same thr: it must contain every candidate
    for (int k : cand) if (a[k] > thr) ok = false;
    if (!ok) {
      if (bad < 5) fprintf(stderr, "trial %d fam %d: cand %zu flagged %d many %d i1 %d i2 %d kmin %d margin %g m1 %g m2 %g m3 %g\n",
                           t, fam, cand.size(), flagged, many, r.i1, r.i2, kmin, margin, r.m1, r.m2, r.m3);
      ++bad;
    }
    flagged_n += flagged; many_n += many;
    extra += (flagged && cand.size() == 1);
  }
  printf("trials %d bad %ld flagged %ld many %ld spurious_flags %ld\n", trials, bad, flagged_n, many_n, extra);
  return bad ? 1 : 0;
}
