// gmm.cu -- K2 (CUDA-core variant): all-pdf diagonal-GMM frame log-likelihoods as the dense contraction
//   ll[t, m] = g_m + sum_d (mu/sigma^2)_md x_td + sum_d (-1/2 sigma^2)_md x_td^2      (m = Gaussian)
//   out[t, j] = logsumexp_{m in pdf j} ll[t, m]
// with fp32 FFMA accumulation in the order g, means term d=0..D-1, variance term d=0..D-1.
//
// Replaces DecodableAmDiagGmmScaled::LogLikelihoodZeroBased / gmm_compute_likes (reference call sites:
// montreal_forced_aligner/alignment/multiprocessing.py:846 (inside GmmAligner), :1415).  Semantics per SURVEY.md A.4
// (Kaldi gmm/decodable-am-diag-gmm.cc, matrix/kaldi-vector.cc LogSumExp incl. the max+log(FLT_EPSILON) cutoff).
//
// This is the exact-order fp32 kernel the tensor-core kernel (gmm_tc.cu) is cross-checked against, and the path
// for models whose shapes the tensor-core kernel does not cover.  Output layout is pdf-major: llT[pdf][ld].
#include "cuda_internal.cuh"

using namespace mfa;

namespace {
constexpr int TM = 128;            // frames per tile
constexpr int TN = MFA_TILE_N;     // Gaussian rows per tile
constexpr int CLD = TN + 1;        // padded C row (conflict-free column walks)
constexpr float kMinLogDiff = -15.9423847198486328125f;  // logf(FLT_EPSILON)

__global__ void __launch_bounds__(256, 1)
gmm_ffma_kernel(const float *__restrict__ feats, int64_t n_rows, int dim, int kdim, const float *__restrict__ W, const float *__restrict__ G,
                const int32_t *__restrict__ tile_pdf0, const int32_t *__restrict__ tile_seg, int n_tiles, float *__restrict__ llT, int64_t ld) {
  extern __shared__ float sm[];
  float *As = sm;                    // [kdim][TM]
  float *Bs = As + kdim * TM;        // [kdim][TN]
  float *Cs = Bs + kdim * TN;        // [TM][CLD]
  __shared__ int seg[TN + 1];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  // A tile: x and x^2, k-major
  for (int i = tid; i < TM * dim; i += 256) {
    int r = i % TM, d = i / TM;
    int64_t row = row0 + r;
    float v = (row < n_rows) ? feats[row * dim + d] : 0.0f;
    As[d * TM + r] = v;
    As[(dim + d) * TM + r] = v * v;
  }
  const int ty = tid >> 4, tx = tid & 15;
  for (int tl = blockIdx.y; tl < n_tiles; tl += gridDim.y) {
    __syncthreads();  // previous iteration's readers of Bs / Cs / seg are done; As visible on the first pass
    const float4 *Wt = (const float4 *)(W + (size_t)tl * kdim * TN);
    for (int i = tid; i < kdim * TN / 4; i += 256) ((float4 *)Bs)[i] = Wt[i];
    for (int i = tid; i <= TN; i += 256) seg[i] = tile_seg[(size_t)tl * (TN + 1) + i];
    __syncthreads();
    float acc[8][8];
    {
      const float *g = G + (size_t)tl * TN;
      float gc[8];
#pragma unroll
      for (int j = 0; j < 4; j++) { gc[j] = g[tx * 4 + j]; gc[4 + j] = g[64 + tx * 4 + j]; }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = gc[j];
    }
#pragma unroll 4
    for (int k = 0; k < kdim; k++) {
      float4 a0 = *(const float4 *)(As + k * TM + ty * 4), a1 = *(const float4 *)(As + k * TM + 64 + ty * 4);
      float4 b0 = *(const float4 *)(Bs + k * TN + tx * 4), b1 = *(const float4 *)(Bs + k * TN + 64 + tx * 4);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int c = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
        Cs[r * CLD + c] = acc[i][j];
      }
    }
    __syncthreads();
    const int p0 = tile_pdf0[tl], np = tile_pdf0[tl + 1] - p0;
    const int r = tid & (TM - 1);
    const int64_t row = row0 + r;
    for (int k = tid >> 7; k < np; k += 2) {
      const int c0 = seg[k], c1 = seg[k + 1];
      const float *c = Cs + r * CLD;
      float mx = -INFINITY;
      for (int j = c0; j < c1; j++) mx = fmaxf(mx, c[j]);
      const float cutoff = mx + kMinLogDiff;
      float s = 0.0f;
      for (int j = c0; j < c1; j++) { float v = c[j]; if (v >= cutoff) s += __expf(v - mx); }
      if (row < ld) llT[(size_t)(p0 + k) * ld + row] = mx + __logf(s);
    }
  }
}
}  // namespace

namespace mfa {
int launch_gmm_ffma(mfa_engine *e, mfa_model *m, const float *d_feats, int64_t n_rows, float *d_llT, int64_t ld) {
  if (n_rows == 0) return MFA_OK;
  if (ld < n_rows) return set_error(MFA_ERR_INVALID, "ld < n_rows");
  MFA_TRY(m->ensure_ffma());
  size_t smem = ((size_t)m->kdim * TM + (size_t)m->kdim * TN + (size_t)TM * CLD) * sizeof(float);
  if (smem > e->smem_optin) return set_error(MFA_ERR_UNSUPPORTED, "GMM tile exceeds shared memory");
  CUDA_TRY(cudaFuncSetAttribute(gmm_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t ftiles = (n_rows + TM - 1) / TM;
  int64_t gsplit = (2 * (int64_t)e->sm_count + ftiles - 1) / ftiles;
  if (gsplit < 1) gsplit = 1;
  if (gsplit > m->n_tiles) gsplit = m->n_tiles;
  if (ftiles > 2147483647LL) return set_error(MFA_ERR_UNSUPPORTED, "too many frames in one launch");
  dim3 grid((unsigned)ftiles, (unsigned)gsplit);
  gmm_ffma_kernel<<<grid, 256, smem, e->stream>>>(d_feats, n_rows, m->dim, m->kdim, m->d_W, m->d_G, m->d_tile_pdf0, m->d_tile_seg, m->n_tiles,
                                                   d_llT, ld);
  e->launches++;
  e->gmm_flops += 2.0 * (2 * m->dim + 1) * (double)m->num_gauss * (double)n_rows;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}
}  // namespace mfa
