// Corpus-side kernels of the semantic-id table: the formats either side of the tokeniser.
//
//   rqb200_sid_dedup_rank   modules/tokenizer/semids.py:94-108: for every corpus row, how many EARLIER rows carry the identical
//                           id tuple (the reference's O(N^2) compare, 90-97 % of its corpus pass), plus the diversity statistics
//                           of train_rqvae.py:276-283 (max duplicates, number of distinct tuples, entropy of the tuple
//                           distribution) from the same pass.
//   rqb200_sid_gather       semids.py:112-146: cached_ids[item_ids] -> [B, S * C] token rows with -1 under the padding mask, and
//                           the matching token_type_ids, in one launch.
//
// Dedup without a sort: the packed tuple (K^L <= 2^26 keys: 24 bits for K = 256, L = 3) addresses a head table; pass 1 threads
// every row onto its key's list with one atomicExch; pass 2 walks the (short) list of the row's key and counts the members with
// a smaller row index.  Work is sum over keys of (group size)^2, i.e. O(N) for the near-unique tables a trained model produces,
// and never worse than the reference's O(N^2).
#include "common.cuh"

#define SID_MAX_KEYS (1ll << 26)

static int64_t sid_key_space(int L, int K) {
  int64_t s = 1;
  for (int l = 0; l < L; ++l) {
    s *= K;
    if (s > SID_MAX_KEYS) return 0;
  }
  return s;
}

extern "C" size_t rqb200_sid_dedup_workspace_bytes(int N, int L, int K) {
  const int64_t keys = sid_key_space(L, K);
  if (keys == 0 || N < 0) return 0;                      // key space too large for a direct table: the caller sorts instead
  return (size_t)(keys + N) * sizeof(int) + 64;
}

__device__ __forceinline__ int64_t sid_pack(const int64_t* row, int L, int K, bool& ok) {
  int64_t key = 0;
  ok = true;
  for (int l = 0; l < L; ++l) {
    const int64_t v = row[l];
    ok = ok && v >= 0 && v < K;
    key = key * K + v;
  }
  return key;
}

__global__ void sid_link_kernel(const int64_t* ids, int N, int L, int K, int* head, int* next) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    bool ok;
    const int64_t key = sid_pack(ids + (int64_t)i * L, L, K, ok);
    next[i] = ok ? atomicExch(&head[key], i) : -2;       // -2: an id outside [0, K): the row is its own group
  }
}

// stats: [0] max rank, [1] distinct tuples; entropy: -sum p log p over distinct tuples (p = group size / N)
__global__ void sid_rank_kernel(const int64_t* ids, int N, int L, int K, const int* head, const int* next, int64_t* rank,
                                int* stats, double* entropy) {
  double ent = 0.0;
  int mx = 0, uniq = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    int r = 0, g = 1;
    if (next[i] != -2) {
      bool ok;
      const int64_t key = sid_pack(ids + (int64_t)i * L, L, K, ok);
      g = 0;
      for (int j = head[key]; j >= 0; j = next[j]) {     // the list holds exactly the rows with this key
        r += (j < i);
        ++g;
      }
    }
    rank[i] = r;
    mx = max(mx, r);
    if (r == 0) {                                        // the earliest row of its group speaks for the group
      ++uniq;
      const double p = (double)g / (double)N;
      ent -= p * log(p);
    }
  }
  ent = warp_sum_d(ent);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    uniq += __shfl_xor_sync(0xffffffffu, uniq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&stats[0], mx);
    atomicAdd(&stats[1], uniq);
    atomicAdd(entropy, ent);
  }
}

extern "C" int rqb200_sid_dedup_rank(const int64_t* ids, int N, int L, int K, int64_t* rank, int* stats, double* entropy,
                                     void* workspace, size_t ws_bytes, void* stream) {
  const int64_t keys = sid_key_space(L, K);
  if (keys == 0) {
    rqb_set_error("sid_dedup_rank: key space K^L = %d^%d exceeds the direct table (2^26 keys)", K, L);
    return RQB_ERR_UNSUPPORTED;
  }
  RQB_CHECK_ARG(N >= 0 && L > 0 && K > 0 && stats && entropy, "sid_dedup_rank: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  RQB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int), st));
  RQB_CUDA(cudaMemsetAsync(entropy, 0, sizeof(double), st));
  if (N == 0) return RQB_OK;
  RQB_CHECK_ARG(ids && rank && workspace, "sid_dedup_rank: null pointer");
  if (ws_bytes < rqb200_sid_dedup_workspace_bytes(N, L, K)) {
    rqb_set_error("sid_dedup_rank: workspace too small");
    return RQB_ERR_WORKSPACE;
  }
  int* head = reinterpret_cast<int*>(workspace);
  int* next = head + keys;
  RQB_CUDA(cudaMemsetAsync(head, 0xFF, (size_t)keys * sizeof(int), st));     // -1 = empty list
  int grid = (N + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  sid_link_kernel<<<grid, 256, 0, st>>>(ids, N, L, K, head, next);
  RQB_LAUNCH_CHECK();
  sid_rank_kernel<<<grid, 256, 0, st>>>(ids, N, L, K, head, next, rank, stats, entropy);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}

// out[b, s * C + c] = mask[b, s] ? cached[item[b, s], c] : -1;   token_type[b, s * C + c] = c     (mask may be null: all valid)
__global__ void sid_gather_kernel(const int64_t* __restrict__ cached, int C, const int64_t* __restrict__ item, int64_t item_stride,
                                  const unsigned char* __restrict__ mask, int64_t mask_stride, int B, int S, int64_t* __restrict__ out,
                                  int64_t* __restrict__ token_type) {
  const int64_t total = (int64_t)B * S * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t bs = e / C;
    const int s = (int)(bs % S);
    const int64_t b = bs / S;
    const bool valid = mask == nullptr || mask[b * mask_stride + s] != 0;
    int64_t v = -1;
    if (valid) v = cached[item[b * item_stride + s] * C + c];
    out[e] = v;
    if (token_type) token_type[e] = c;
  }
}

extern "C" int rqb200_sid_gather(const int64_t* cached_ids, int64_t n_corpus, int C, const int64_t* item_ids, int64_t item_stride,
                                 const unsigned char* seq_mask, int64_t mask_stride, int B, int S, int64_t* out,
                                 int64_t* token_type, void* stream) {
  RQB_CHECK_ARG(B >= 0 && S >= 0 && C > 0 && n_corpus >= 0, "sid_gather: bad shape");
  if ((int64_t)B * S == 0) return RQB_OK;
  RQB_CHECK_ARG(cached_ids && item_ids && out, "sid_gather: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)B * S * C;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  sid_gather_kernel<<<grid, 256, 0, st>>>(cached_ids, C, item_ids, item_stride, seq_mask, mask_stride, B, S, out, token_type);
  RQB_LAUNCH_CHECK();
  return RQB_OK;
}
