// Python-free harness around the C ABI's tensor-core tokeniser (include/rqb200.h: rqb200_tokenize_tc_prepare / _run):
// seeded synthetic unit-norm rows and live codebooks, one run whose ids go to a file, then `iters` event-timed runs.
// Two builds of the library (e.g. -DTX_R=96 vs the default tile shape) are compared by running this binary once per build
// (RQB200_LIB=path/to/other.so) and `cmp`-ing the id files.
// A fresh GPU box spends about a minute importing torch; this starts in milliseconds, which matters when GPU time is short.
// build: nvcc -O2 -std=c++17 -o tools/bin/tc_native_check tools/tc_native_check.cu -ldl      (run from the repo root)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 2; } } while (0)

typedef size_t (*StateBytesFn)(int, int, int);
typedef int (*PrepareFn)(const float* const*, int, int, int, void*, size_t, void*);
typedef int (*RunFn)(const float*, int64_t, int, const void*, int, int, int, int64_t*, int*, void*);
typedef const char* (*ErrFn)();

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline float urand() {   // xorshift64*, uniform in [0,1)
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) / 16777216.0f;
}
static inline float grand() { return sqrtf(-2.f * logf(urand() + 1e-12f)) * cosf(6.2831853f * urand()); }

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 65536, D = argc > 2 ? atoi(argv[2]) : 768, L = argc > 3 ? atoi(argv[3]) : 3;
  const int iters = argc > 4 ? atoi(argv[4]) : 20;
  const char* out = argc > 5 ? argv[5] : nullptr;
  const int K = 256;
  const char* libpath = getenv("RQB200_LIB") ? getenv("RQB200_LIB") : "rq_vae_recommender_b200/librqb200.so";
  void* lib = dlopen(libpath, RTLD_NOW);
  if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
  auto state_bytes = (StateBytesFn)dlsym(lib, "rqb200_tokenize_tc_state_bytes");
  auto prepare = (PrepareFn)dlsym(lib, "rqb200_tokenize_tc_prepare");
  auto run = (RunFn)dlsym(lib, "rqb200_tokenize_tc_run");
  auto last_error = (ErrFn)dlsym(lib, "rqb200_last_error");
  if (!state_bytes || !prepare || !run || !last_error) { printf("missing symbol\n"); return 2; }

  std::vector<float> x((size_t)B * D);
  // the inputs are cached in /tmp between runs of a session (the generator costs seconds at 65 536 rows, a GPU box is paid by the minute)
  char cache[256];
  snprintf(cache, sizeof cache, "/tmp/tc_native_x_%d_%d.bin", B, D);
  bool cached = false;
  if (FILE* f = fopen(cache, "rb")) { cached = fread(x.data(), 4, x.size(), f) == x.size(); fclose(f); }
  if (!cached) {
  for (int b = 0; b < B; ++b) {
    double n2 = 0;
    for (int d = 0; d < D; ++d) { const float v = grand(); x[(size_t)b * D + d] = v; n2 += (double)v * v; }
    const float inv = (float)(1.0 / sqrt(n2));
    for (int d = 0; d < D; ++d) x[(size_t)b * D + d] *= inv;
  }
  if (FILE* f = fopen(cache, "wb")) { fwrite(x.data(), 4, x.size(), f); fclose(f); }
  }
  rng_state = 0x1234567887654321ull ^ (uint64_t)B;      // the codebooks below are drawn from a stream that does not depend on the cache
  // level 0: rows of x plus noise (all codes live); level l: residual-sized random directions
  std::vector<std::vector<float>> cbs(L, std::vector<float>((size_t)K * D));
  for (int l = 0; l < L; ++l)
    for (int k = 0; k < K; ++k) {
      const size_t src = ((size_t)k * 2654435761u + l * 97u) % (size_t)B;
      const float s = l == 0 ? 1.f : 0.f, noise = (l == 0 ? 0.3f : 0.7f / (float)l) / sqrtf((float)D);
      for (int d = 0; d < D; ++d) cbs[l][(size_t)k * D + d] = s * x[src * D + d] + noise * grand();
    }
  float* dx; int64_t* dids; int* dstats; void* dstate;
  std::vector<float*> dcb(L);
  CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dids, (size_t)B * L * 8)); CK(cudaMalloc(&dstats, 64 * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dids, 0xff, (size_t)B * L * 8)); CK(cudaMemset(dstats, 0, 64 * 4));
  for (int l = 0; l < L; ++l) {
    CK(cudaMalloc(&dcb[l], (size_t)K * D * 4));
    CK(cudaMemcpy(dcb[l], cbs[l].data(), (size_t)K * D * 4, cudaMemcpyHostToDevice));
  }
  const size_t sb = state_bytes(D, K, L);
  if (!sb) { printf("shape not supported\n"); return 2; }
  CK(cudaMalloc(&dstate, sb));
  if (prepare(dcb.data(), D, K, L, dstate, sb, nullptr)) { printf("prepare failed: %s\n", last_error()); return 1; }
  if (run(dx, D, B, dstate, D, K, L, dids, dstats, nullptr)) { printf("run failed: %s\n", last_error()); return 1; }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("first run: %s\n", cudaGetErrorString(se)); return 1; }
  std::vector<int64_t> ids((size_t)B * L);
  int stats[4];
  CK(cudaMemcpy(ids.data(), dids, ids.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(stats, dstats, 16, cudaMemcpyDeviceToHost));
  long bad = 0; uint64_t h = 1469598103934665603ull;
  for (int64_t v : ids) { if (v < 0 || v >= K) ++bad; h = (h ^ (uint64_t)v) * 1099511628211ull; }
  if (out) { FILE* f = fopen(out, "wb"); if (f) { fwrite(ids.data(), 8, ids.size(), f); fclose(f); } }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) run(dx, D, B, dstate, D, K, L, dids, nullptr, nullptr);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) run(dx, D, B, dstate, D, K, L, dids, nullptr, nullptr);
  CK(cudaEventRecord(e1));
  se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("timed runs: %s\n", cudaGetErrorString(se)); return 1; }
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  const char* vpf = getenv("RQB200_TC_PREFETCH");
  printf("B=%d D=%d L=%d lib=%s PREFETCH=%s: ids out of range %ld, fnv %016llx, re-ranked rows %d cands %d many %d, %.4f ms/run (%d runs), %.1f M items/s\n",
         B, D, L, libpath, vpf ? vpf : "-", bad, (unsigned long long)h, stats[0], stats[1], stats[2], ms / iters, iters,
         B / (ms / iters) * 1e-3);
  // RQB200_TC_TRACE=1: one more run with the event timeline of CTA 0 switched on (stats[3] = stats[4] = 1, >= 4096 ints; the
  // records are (tag << 56 | payload << 48 | clock) at ((long long*)(stats + 128))[role * 256 ...], see TC_EV in tc_common.cuh)
  const char* vtr = getenv("RQB200_TC_TRACE");
  if (vtr && vtr[0] == '1') {
    int* dtr; CK(cudaMalloc(&dtr, 4096 * 4)); CK(cudaMemset(dtr, 0, 4096 * 4));   // [0,128) counters | 4 x 256 events | 24 x 8 x 5 phase clocks
    const int on[2] = {1, 1};
    CK(cudaMemcpy(dtr + 3, on, 8, cudaMemcpyHostToDevice));
    if (run(dx, D, B, dstate, D, K, L, dids, dtr, nullptr)) { printf("traced run failed: %s\n", last_error()); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<int> tr(4096);
    CK(cudaMemcpy(tr.data(), dtr, 4096 * 4, cudaMemcpyDeviceToHost));
    // cycle accounting of the traced instantiation (64-bit accumulators from stats[8]; slot 12 = number of MMA-issuing CTAs)
    {
      const long long* acc = reinterpret_cast<const long long*>(tr.data() + 8);
      static const struct { int slot; const char* name; } names[] = {
          {0, "mma wait t_empty"}, {1, "mma wait a_full"}, {2, "mma wait b_full"}, {3, "mma total"},
          {9, "conv wait A slot (a_empty)"}, {20, "conv wait x (loads landed / x_full)"}, {21, "conv convert+store"}, {10, "conv total"},
          {4, "epi0 wait t_full"}, {5, "epi0 scan"}, {6, "epi0 wait partner warps"}, {7, "epi0 merge+many+rerank"}, {8, "epi0 total"},
          {13, "epi1 wait t_full"}, {14, "epi1 scan"}, {15, "epi1 wait"}, {16, "epi1 other"}, {17, "epi1 total"}};
      const double nb = acc[12] > 0 ? (double)acc[12] : 1.0;
      printf("cycle accounting per MMA-issuing CTA (%lld of them):\n", acc[12]);
      for (const auto& n : names) printf("  %-38s %12.0f\n", n.name, (double)acc[n.slot] / nb);
    }
    const long long* ev = reinterpret_cast<const long long*>(tr.data() + 128);
    struct Rec { long long clk; int role, tag, pay; };
    std::vector<Rec> recs;
    for (int r = 0; r < 4; ++r)
      for (int i = 0; i < 256; ++i) {
        const long long e = ev[r * 256 + i];
        if (e) recs.push_back(Rec{e & 0xffffffffffffLL, r, (int)((e >> 56) & 0xff), (int)((e >> 48) & 0xff)});
      }
    for (size_t i = 1; i < recs.size(); ++i)      // insertion sort by clock (a few hundred records)
      for (size_t j = i; j > 0 && recs[j].clk < recs[j - 1].clk; --j) { Rec t = recs[j]; recs[j] = recs[j - 1]; recs[j - 1] = t; }
    static const char* role_name[4] = {"MMA ", "CONV", "EPI0", "EPI1"};
    static const char* tag_name[4][6] = {
        {"", "level start (t_empty ok)", "a_full ok, chunk step", "level issued", "", ""},
        {"", "a_empty/x_full ok, chunk step", "chunk converted+arrived", "", "", ""},
        {"", "t_full ok", "scan end", "merged / verdict out (tcx: exchange complete)", "tmem released (tcx: level done, ids out)", "level done (re-rank, id out)"},
        {"", "t_full ok", "scan end", "id received (tcx: exchange complete)", "tmem released (tcx: level done, ids out)", "level done (re-rank, id out)"}};
    printf("timeline of CTA 0 (cycles since first event; payload = tile_index*16 + level-or-step), %zu events\n", recs.size());
    {   // rq_tcx_kernel: per-warp phase clocks of the epilogue (CTA 0): scan end, exchange complete, queue complete, re-rank done, all done
      const long long* ph = reinterpret_cast<const long long*>(tr.data() + 2176);
      if (ph[0]) {
        printf("epilogue phases of CTA 0, cycles since the step's earliest scan end; per warp: scan_end  +x_full  +phase1/sync  +re-rank  +sync\n");
        for (int g = 0; g < 24; ++g) {
          long long t0 = 0;
          for (int e = 0; e < 8; ++e) { const long long v = ph[(g * 8 + e) * 5]; if (v && (!t0 || v < t0)) t0 = v; }
          if (!t0) break;
          printf("step %2d:", g);
          for (int e = 0; e < 8; ++e) {
            const long long* q = ph + (g * 8 + e) * 5;
            printf("  w%d %5lld %5lld %5lld %5lld %5lld |", e, q[0] - t0, q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3]);
          }
          printf("\n");
        }
      }
    }
    for (const Rec& q : recs) {
      const bool step = (q.role == 0 && q.tag == 2) || q.role == 1;
      if (step && (q.pay & 15) != 0 && (q.pay & 15) != (D / 64 - 1)) continue;      // chunk steps: first and last only
      printf("%9lld  %*s%s %-32s tile %d %s %d\n", q.clk - recs[0].clk, 4 * q.role, "", role_name[q.role],
             q.tag < 6 ? tag_name[q.role][q.tag] : "?", q.pay >> 4, step ? "step" : "lvl", q.pay & 15);
    }
  }
  return bad ? 1 : 0;
}
