// accstats.cu -- K4: GMM accumulator statistics for the align -> acc-stats -> update training loop.
//
// Replaces GmmStatsAccumulator.accumulate_stats / AccumAmDiagGmm.acc_stats + TransitionModel.acc_stats
// (reference call sites: montreal_forced_aligner/alignment/multiprocessing.py:652-666,
// acoustic_modeling/monophone.py:114-120).  Semantics per SURVEY.md A.8 (Kaldi gmmbin/gmm-acc-stats-ali.cc,
// gmm/mle-diag-gmm.cc AccumulateFromPosteriors, gmm/diag-gmm.cc ComponentPosteriors): fp32 component
// log-likelihoods and soft-max posteriors of the ALIGNED pdf only, f64 accumulators
//   occ[m] += g, mean_acc[m,:] += g x, var_acc[m,:] += g x^2, trans_acc[tid] += 1, tot_like += logsumexp.
// One warp per frame; lanes span the feature dimension; per-component reductions by warp shuffle; f64 atomics
// (red.global.add.f64) into the accumulator block that the host layer later all-reduces with NCCL.
#include "cuda_internal.cuh"

using namespace mfa;

namespace {
constexpr int AW = 8;  // warps per CTA

__global__ void __launch_bounds__(AW * 32)
acc_stats_kernel(const float *__restrict__ feats, const int32_t *__restrict__ ali, int64_t n_frames, int dim, int num_gauss, int num_tids,
                 const int32_t *__restrict__ pdf_off, const int32_t *__restrict__ tid2pdf, const float *__restrict__ gconsts,
                 const float *__restrict__ miv, const float *__restrict__ iv, double *__restrict__ acc) {
  __shared__ float post[AW][MFA_TILE_N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *occ = acc, *mean = acc + num_gauss, *var = mean + (size_t)num_gauss * dim, *trans = var + (size_t)num_gauss * dim;
  double *tot = trans + (num_tids + 1);
  double like = 0.0, frames = 0.0;
  const int64_t stride = (int64_t)gridDim.x * AW;
  for (int64_t f = (int64_t)blockIdx.x * AW + warp; f < n_frames; f += stride) {
    const int tid = ali[f];
    if (tid <= 0 || tid > num_tids) continue;
    const int pdf = tid2pdf[tid];
    const int m0 = pdf_off[pdf], nm = pdf_off[pdf + 1] - m0;
    const float x0 = lane < dim ? feats[f * dim + lane] : 0.0f;
    const float x1 = lane + 32 < dim ? feats[f * dim + lane + 32] : 0.0f;
    float mx = -INFINITY;
    for (int m = 0; m < nm; m++) {
      const float *a = miv + (size_t)(m0 + m) * dim, *b = iv + (size_t)(m0 + m) * dim;
      float d1 = 0.0f, d2 = 0.0f;
      if (lane < dim) { d1 = a[lane] * x0; d2 = b[lane] * (x0 * x0); }
      if (lane + 32 < dim) { d1 += a[lane + 32] * x1; d2 += b[lane + 32] * (x1 * x1); }
#pragma unroll
      for (int o = 16; o; o >>= 1) { d1 += __shfl_xor_sync(0xffffffffu, d1, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
      float v = gconsts[m0 + m] + d1;
      v = v + (-0.5f) * d2;
      if (lane == 0) post[warp][m] = v;
      mx = fmaxf(mx, v);
    }
    __syncwarp();
    float sum = 0.0f;
    for (int m = lane; m < nm; m += 32) { float ev = expf(post[warp][m] - mx); post[warp][m] = ev; sum += ev; }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
    like += (double)(mx + logf(sum));
    frames += 1.0;
    if (lane == 0) atomicAdd(&trans[tid], 1.0);
    const double xd0 = x0, xd1 = x1;
    for (int m = 0; m < nm; m++) {
      const double g = (double)(post[warp][m] * inv);
      if (lane == 0) atomicAdd(&occ[m0 + m], g);
      if (lane < dim) { atomicAdd(&mean[(size_t)(m0 + m) * dim + lane], g * xd0); atomicAdd(&var[(size_t)(m0 + m) * dim + lane], g * (xd0 * xd0)); }
      if (lane + 32 < dim) { atomicAdd(&mean[(size_t)(m0 + m) * dim + lane + 32], g * xd1); atomicAdd(&var[(size_t)(m0 + m) * dim + lane + 32], g * (xd1 * xd1)); }
    }
    __syncwarp();
  }
  if (lane == 0 && frames > 0.0) { atomicAdd(&tot[0], like); atomicAdd(&tot[1], frames); }
}
}  // namespace

extern "C" int64_t mfa_acc_size(const mfa_model *m) {
  if (!m) return 0;
  return (int64_t)m->num_gauss * (1 + 2 * (int64_t)m->dim) + m->num_tids + 1 + 2;
}

extern "C" int mfa_acc_zero(mfa_engine *e, mfa_model *m) {
  if (!e || !m) return set_error(MFA_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  size_t bytes = (size_t)mfa_acc_size(m) * sizeof(double);
  if (!m->d_acc) CUDA_TRY(cudaMalloc((void **)&m->d_acc, bytes));
  CUDA_TRY(cudaMemsetAsync(m->d_acc, 0, bytes, e->stream));
  return MFA_OK;
}

extern "C" double *mfa_acc_device_ptr(mfa_engine *e, mfa_model *m) { (void)e; return m ? m->d_acc : nullptr; }

extern "C" int mfa_acc_read(mfa_engine *e, mfa_model *m, double *host_out) {
  if (!e || !m || !host_out || !m->d_acc) return set_error(MFA_ERR_INVALID, "no accumulators");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaMemcpyAsync(host_out, m->d_acc, (size_t)mfa_acc_size(m) * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

namespace mfa {
int launch_acc_stats(mfa_engine *e, mfa_model *m, const float *d_feats, const int32_t *d_ali, int64_t n_frames) {
  if (n_frames == 0) return MFA_OK;
  if (!m->d_acc) MFA_TRY(mfa_acc_zero(e, m));
  int64_t blocks = (n_frames + AW - 1) / AW;
  int64_t cap = (int64_t)e->sm_count * 16;
  if (blocks > cap) blocks = cap;
  acc_stats_kernel<<<(unsigned)blocks, AW * 32, 0, e->stream>>>(d_feats, d_ali, n_frames, m->dim, m->num_gauss, m->num_tids, m->d_pdf_off,
                                                                 m->d_tid2pdf, m->d_gconsts, m->d_miv, m->d_iv, m->d_acc);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}
}  // namespace mfa
