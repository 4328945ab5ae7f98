// gmm_tc.cu -- K2 on tcgen05 tensor cores (placeholder: forwards to the CUDA-core kernel until the tcgen05 kernel lands).
#include "cuda_internal.cuh"
namespace mfa {
int launch_gmm_tc(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t n_rows, float *d_llT, int64_t ld) {
  return launch_gmm_ffma(e, m, d_feats, n_rows, d_llT, ld);
}
}  // namespace mfa
