// Tensor-core (tcgen05) tokeniser: candidate filter on fp16 tensor cores + exact fp32 re-rank.
// (placeholder translation unit while the kernel is being brought up: reports "unsupported" for every shape so
//  callers use the exact CUDA-core kernel of rq_simt.cu)
#include "common.cuh"

extern "C" size_t rqb200_tokenize_tc_state_bytes(int D, int K, int L) { return 0; }
extern "C" int rqb200_tokenize_tc_supported(int D, int K, int L) { return 0; }
extern "C" int rqb200_tokenize_tc_prepare(const float* const* codebooks, int D, int K, int L, void* state,
                                          size_t state_bytes, void* stream) {
  rqb_set_error("tokenize_tc: shape D=%d K=%d L=%d not supported", D, K, L);
  return RQB_ERR_UNSUPPORTED;
}
extern "C" int rqb200_tokenize_tc_run(const float* x, int64_t ldx, int B, const void* state, int D, int K, int L,
                                      int64_t* ids, int* stats, void* stream) {
  rqb_set_error("tokenize_tc: shape D=%d K=%d L=%d not supported", D, K, L);
  return RQB_ERR_UNSUPPORTED;
}
