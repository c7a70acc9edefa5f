// How many bytes per clock can the SMs pull out of L2 with bulk copies (TMA), and does it matter whether all SMs stream the SAME
// lines (the tokeniser's codebook blocks: 1.18 MB read by every CTA for every tile) or private ones?  DESIGN.md 5.2d: the
// tokeniser kernels all land at 16-20 B/clk/SM of L2->SM traffic whatever they keep in flight; this probe measures the ceiling.
//   grid = #SMs, one CTA per SM (227 KB of dynamic shared memory requested so that nothing co-resides), a ring of `depth`
//   stages of `chunk` bytes, one elected thread issues cp.async.bulk global->shared and waits on the stage's mbarrier;
//   region: `shared` = every CTA walks the same `span` bytes; `private` = CTA b walks [b * span, (b+1) * span)
//   (span is small enough to stay in the 126 MB L2: the first pass warms it, only later passes are timed)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/l2_stream_probe tools/l2_stream_probe.cu
// run:   tools/bin/l2_stream_probe            (prints a table: pattern x chunk x depth -> B/clk/SM, TB/s)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

// passes over the CTA's region; pass 0 is the warm-up (untimed)
__global__ void __launch_bounds__(128, 1) stream_kernel(const unsigned char* base, size_t span, size_t cta_stride, int chunk, int depth,
                                                        int passes, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // [depth]
  unsigned char* ring = smem + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned char* src = base + (size_t)blockIdx.x * cta_stride;
  const int per_pass = (int)(span / chunk);
  long long t0 = 0;
  uint32_t issued = 0, waited = 0;
  const int total = per_pass * passes;
  // prologue: fill the ring
  for (; issued < (uint32_t)depth && issued < (uint32_t)total; ++issued) {
    mbar_expect(&bars[issued % depth], chunk);
    bulk_g2s(ring + (size_t)(issued % depth) * chunk, src + (size_t)(issued % per_pass) * chunk, chunk, &bars[issued % depth]);
  }
  for (; waited < (uint32_t)total; ++waited) {
    if (waited == (uint32_t)per_pass) t0 = clock64();            // pass 0 done: start timing
    const uint32_t st = waited % depth;
    while (!mbar_try(&bars[st], (waited / depth) & 1)) {}
    if (issued < (uint32_t)total) {                              // refill the stage just drained
      mbar_expect(&bars[st], chunk);
      bulk_g2s(ring + (size_t)st * chunk, src + (size_t)(issued % per_pass) * chunk, chunk, &bars[st]);
      ++issued;
    }
  }
  cycles[blockIdx.x] = clock64() - t0;
}

// ---- second question: how many uncoalesced 32-byte gather requests per clock does one SM's L1TEX sustain?  (The epilogue's Gram
// gathers: every lane its own 128-byte line, LDG.E.256.)  `warps` warps per CTA each issue `n` rounds of `batch` independent
// 256-bit loads; lane l of warp w reads sector (round-dependent) of row (l + 32 w + 97 * round) % rows of a [rows][1 KB] table
// that stays in L2 (768 KB, the size of the three Gram tables).
__global__ void __launch_bounds__(256, 1) gather_kernel(const float* table, int rows, int rounds, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];        // only there to keep the carve-out the tokeniser runs with
  (void)smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    float4 lo[4], hi[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {                                // 4 independent 256-bit loads in flight per lane, like the scan at level 2
      const int row = (lane + 32 * warp + 97 * (r * 4 + b) + 13 * (int)blockIdx.x) % rows;
      const float* q = table + (size_t)row * 256 + ((r * 4 + b) & 31) * 8;
      asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=f"(lo[b].x), "=f"(lo[b].y), "=f"(lo[b].z), "=f"(lo[b].w), "=f"(hi[b].x), "=f"(hi[b].y), "=f"(hi[b].z), "=f"(hi[b].w)
                   : "l"(q));
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) acc += lo[b].x + hi[b].w;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const size_t span = 1179648;                       // the tokeniser's fp16 codebook blob at the north-star shape (3 x 256 x 768 x 2)
  unsigned char* buf; long long* dcyc;
  CK(cudaMalloc(&buf, span * (size_t)sms));          // 175 MB: private regions; together they exceed L2, so "private" is partly an HBM test
  CK(cudaMemset(buf, 1, span * (size_t)sms));
  CK(cudaMalloc(&dcyc, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  printf("%d SMs, clock rate attribute %d kHz; region %zu bytes per CTA; rows: pattern chunk depth -> in flight, B/clk/SM (slowest / mean CTA), TB/s at the attribute clock\n",
         sms, khz, span);
  const int passes = 9;                              // 1 warm-up + 8 timed
  for (int pattern = 0; pattern < 3; ++pattern)      // 0 shared by all, 1 private per CTA (L2 + HBM), 2 private but only 64 CTAs' worth (fits L2)
    for (int chunk : {4096, 16384, 32768})
      for (int depth : {2, 4, 6}) {
        if ((size_t)chunk * depth + 1024 > 232448) continue;
        const size_t stride = pattern == 0 ? 0 : span;
        const int grid = pattern == 2 ? 64 : sms;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        stream_kernel<<<grid, 128, 232448>>>(buf, span, stride, chunk, depth, passes, dcyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long cyc[256]; CK(cudaMemcpy(cyc, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0; double sum = 0;
        for (int i = 0; i < grid; ++i) { mx = cyc[i] > mx ? cyc[i] : mx; sum += (double)cyc[i]; }
        const double bytes = (double)span * (passes - 1);
        const double bpc_slow = bytes / (double)mx, bpc_mean = bytes / (sum / grid);
        printf("%-8s chunk %5d depth %d  in flight %6d B  %6.1f / %6.1f B/clk/SM   %6.2f TB/s (%d CTAs, %.3f ms incl. warm-up)\n",
               pattern == 0 ? "shared" : pattern == 1 ? "private" : "priv64", chunk, depth, chunk * depth, bpc_slow, bpc_mean,
               bpc_mean * grid * (double)khz * 1e3 / 1e12, grid, ms);
      }
  // ---- gather request rate
  {
    const int rows = 768;                                        // 768 rows x 1 KB = the three Gram tables
    float *tab, *sink; CK(cudaMalloc(&tab, (size_t)rows * 1024)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(tab, 0, (size_t)rows * 1024));
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int warps : {1, 2, 4, 8}) {
      const int rounds = 4096;
      gather_kernel<<<sms, warps * 32, 200 * 1024>>>(tab, rows, 64, dcyc, sink);       // warm L2
      gather_kernel<<<sms, warps * 32, 200 * 1024>>>(tab, rows, rounds, dcyc, sink);
      CK(cudaDeviceSynchronize());
      long long cyc[256]; CK(cudaMemcpy(cyc, dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      double sum = 0; for (int i = 0; i < sms; ++i) sum += (double)cyc[i];
      const double req = (double)rounds * 4 * 32 * warps;        // 32-byte sector requests per SM
      printf("gather: %d warps/SM, 4 x LDG.256 in flight per lane, every lane its own line: %.2f requests/clk/SM = %.1f B/clk/SM\n",
             warps, req / (sum / sms), 32.0 * req / (sum / sms));
    }
  }
  return 0;
}
