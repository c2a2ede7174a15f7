// K2 (throughput mode) — tcgen05 / TMEM implicit-GEMM classifier.  Placeholder until the tensor-core path lands.
#include "ss_common.cuh"

namespace ss {

int tc_create(ss_ctx* ctx, const float* blob_host_payload) {
  (void)ctx; (void)blob_host_payload;
  return SS_OK;
}

void tc_destroy(ss_ctx* ctx) { (void)ctx; }

int classify_bf16(ss_ctx* ctx, const float* mel, int n_windows, float* logits, float* spec_out, cudaStream_t st) {
  (void)ctx; (void)mel; (void)n_windows; (void)logits; (void)spec_out; (void)st;
  set_error("SS_MODE_BF16 is not available in this build");
  return SS_E_ARG;
}

}  // namespace ss
