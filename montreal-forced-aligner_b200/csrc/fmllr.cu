// fmllr.cu -- K5: per-speaker fMLLR statistics (row N2 of SURVEY.md section 8f: the stage between the two alignment passes).
//
// Replaces the accumulation half of kalpy FmllrComputer.export_transforms, reached from CalcFmllrFunction._run
// (montreal_forced_aligner/corpus/features.py:460-548; options :759-766).  Semantics = Kaldi gmmbin/gmm-est-fmllr(-gpost).cc +
// transform/fmllr-diag-gmm.cc (FmllrDiagGmmAccs::AccumulateFromPosteriors, CommitSingleFrameStats) + WeightSilencePost:
//   per frame t with aligned pdf j and weight w (silence_weight for silence phones):
//     p_m = w * softmax_m(component log-likelihoods of j under the POSTERIOR model)           (fp32)
//     a = sum_m p_m (mu/sigma^2)_m,  b = sum_m p_m (1/sigma^2)_m   under the STATISTICS model    (fp32)
//     beta += sum_m p_m;  K += a xi^T;  G_d += b_d xi xi^T   with xi = [x; 1]                   (f64)
// Two kernels:
//   fmllr_frame_kernel  one warp per frame, lanes over the feature dimension: writes a | b (fp32) and the frame's count.
//   fmllr_accum_kernel  one CTA per (speaker, tile of 8 rows d, frame split): every thread owns up to NIJ entries (i, j) of the
//                       packed lower triangle x 8 rows of f64 sums in registers -- G_d is the GEMM  B^T[8 x T] * Z[T x 861] with
//                       Z_t = vec(xi xi^T) formed on the fly from the frame staged in shared memory, never materialised -- and
//                       flushes once (f64 red; plain sums when a speaker is not split).  D = 40: 147 G DFMA for 10 h of audio.
// The row-by-row transform update itself (40 iterations x D rows of small dense solves per speaker) runs on the host in f64
// (fmllr.py); it is O(speakers), not O(frames).
#include <algorithm>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {
#ifndef MFA_FMLLR_DT
#define MFA_FMLLR_DT 8
#endif
constexpr int FW = 8;     // warps per CTA in the frame kernel
constexpr int FB = 32;    // frames per staged batch in the accumulation kernel
constexpr int DT = MFA_FMLLR_DT;     // rows d per CTA
constexpr int ANT = 256;  // threads per CTA in the accumulation kernel

__global__ void __launch_bounds__(FW * 32)
fmllr_frame_kernel(const float *__restrict__ feats, const int32_t *__restrict__ ali, int64_t n_frames, int dim, int num_tids,
                   const int32_t *__restrict__ pdf_off, const int32_t *__restrict__ tid2pdf, const float *__restrict__ tid_weight,
                   const float *__restrict__ p_gconsts, const float *__restrict__ p_miv, const float *__restrict__ p_iv,
                   const float *__restrict__ s_miv, const float *__restrict__ s_iv, float *__restrict__ ab, float *__restrict__ cnt) {
  __shared__ float post[FW][MFA_TILE_N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * FW;
  for (int64_t f = (int64_t)blockIdx.x * FW + warp; f < n_frames; f += stride) {
    const int tid = ali[f];
    float w = 0.0f;
    if (tid > 0 && tid <= num_tids) w = tid_weight ? tid_weight[tid] : 1.0f;
    if (w == 0.0f) { if (lane == 0) cnt[f] = 0.0f; continue; }
    const int pdf = tid2pdf[tid];
    const int m0 = pdf_off[pdf], nm = pdf_off[pdf + 1] - m0;
    const float x0 = lane < dim ? feats[f * dim + lane] : 0.0f;
    const float x1 = lane + 32 < dim ? feats[f * dim + lane + 32] : 0.0f;
    float mx = -INFINITY;
    for (int m = 0; m < nm; m++) {
      const float *a = p_miv + (size_t)(m0 + m) * dim, *b = p_iv + (size_t)(m0 + m) * dim;
      float d1 = 0.0f, d2 = 0.0f;
      if (lane < dim) { d1 = a[lane] * x0; d2 = b[lane] * (x0 * x0); }
      if (lane + 32 < dim) { d1 += a[lane + 32] * x1; d2 += b[lane + 32] * (x1 * x1); }
#pragma unroll
      for (int o = 16; o; o >>= 1) { d1 += __shfl_xor_sync(0xffffffffu, d1, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
      float v = p_gconsts[m0 + m] + d1;
      v = v + (-0.5f) * d2;
      if (lane == 0) post[warp][m] = v;
      mx = fmaxf(mx, v);
    }
    __syncwarp();
    float sum = 0.0f;
    for (int m = lane; m < nm; m += 32) { const float ev = expf(post[warp][m] - mx); post[warp][m] = ev; sum += ev; }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
    float a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
    double c = 0.0;
    for (int m = 0; m < nm; m++) {
      const float p = post[warp][m] * inv * w;
      c += (double)p;
      const float *a = s_miv + (size_t)(m0 + m) * dim, *b = s_iv + (size_t)(m0 + m) * dim;
      if (lane < dim) { a0 += a[lane] * p; b0 += b[lane] * p; }
      if (lane + 32 < dim) { a1 += a[lane + 32] * p; b1 += b[lane + 32] * p; }
    }
    float *o = ab + (size_t)f * 2 * dim;
    if (lane < dim) { o[lane] = a0; o[dim + lane] = b0; }
    if (lane + 32 < dim) { o[lane + 32] = a1; o[dim + lane + 32] = b1; }
    if (lane == 0) cnt[f] = (float)c;
    __syncwarp();
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stats per speaker (doubles): beta | K[D][D+1] | G[D][NP], NP = (D+1)(D+2)/2 (row-major lower triangle)
template <int NIJ>
__global__ void __launch_bounds__(ANT, NIJ <= 4 ? (DT <= 4 ? 3 : 2) : 1)
fmllr_accum_kernel(const float *__restrict__ feats, const float *__restrict__ ab, const float *__restrict__ cnt, int dim,
                   const int64_t *__restrict__ frame_off, const int32_t *__restrict__ spk_utt_off, const int32_t *__restrict__ spk_utts,
                   int n_dtile, int n_split, double *__restrict__ stats, int64_t stats_stride) {
  __shared__ __align__(16) double s_xp[FB][66];   // row stride XS below
  __shared__ __align__(16) double s_b[FB][DT], s_a[FB][DT];
  __shared__ float s_c[FB];
  const int t = threadIdx.x;
  const int D1 = dim + 1, NP = D1 * (D1 + 1) / 2;
  int bid = blockIdx.x;
  const int split = bid % n_split; bid /= n_split;
  const int dtile = bid % n_dtile;
  const int spk = bid / n_dtile;
  const int d0 = dtile * DT;
  int pi[NIJ], pj[NIJ];
#pragma unroll
  for (int k = 0; k < NIJ; k++) {
    const int ij = t + k * ANT;
    int i = (int)((sqrtf(8.0f * (float)ij + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= ij) i++;
    while (i * (i + 1) / 2 > ij) i--;
    pi[k] = ij < NP ? i : 0;
    pj[k] = ij < NP ? ij - i * (i + 1) / 2 : 0;
  }
  double g[NIJ][DT];
#pragma unroll
  for (int k = 0; k < NIJ; k++)
#pragma unroll
    for (int d = 0; d < DT; d++) g[k][d] = 0.0;
  double kacc[2] = {0.0, 0.0};
  const int kd0 = t / D1, kj0 = t - kd0 * D1, kd1 = (t + ANT) / D1, kj1 = (t + ANT) - kd1 * D1;
  const bool k0ok = t < DT * D1, k1ok = t + ANT < DT * D1;
  constexpr int XS = 66;   // row stride of s_xp (doubles)
  const double *xi[NIJ], *xj[NIJ];
#pragma unroll
  for (int k = 0; k < NIJ; k++) { xi[k] = &s_xp[0][pi[k]]; xj[k] = &s_xp[0][pj[k]]; }
  // K: the (row, column) pairs this thread owns, clamped to a valid address when it owns none (the sum is then never written)
  const double *ka0 = &s_a[0][k0ok ? kd0 : 0], *kx0 = &s_xp[0][k0ok ? kj0 : 0], *ka1 = &s_a[0][k1ok ? kd1 : 0], *kx1 = &s_xp[0][k1ok ? kj1 : 0];
  double beta = 0.0;
  // Batches of FB frames of this speaker, every n_split-th one for this CTA.  The next batch's global loads are issued into
  // registers before the current batch is consumed (the f64 accumulation hides their latency), staged to shared memory as doubles.
  const int xc = t & 63, xr = t >> 6;          // x: column xc of rows xr, xr+4, ..., xr+28
  const int af = t / DT, ad = t % DT;          // a / b: frame af (< FB for the first FB * DT threads), row d0 + ad
  int ui = spk_utt_off[spk];
  const int ui_end = spk_utt_off[spk + 1];
  int64_t fb = 0, f1 = 0;
  int batch = 0;
  bool have_u = false;
  auto next_batch = [&](int64_t &o_fb, int &o_n) -> bool {   // uniform across the CTA
    for (;;) {
      if (!have_u) {
        if (ui >= ui_end) return false;
        const int u = spk_utts[ui];
        fb = frame_off[u]; f1 = frame_off[u + 1]; have_u = true;
      }
      if (fb >= f1) { have_u = false; ui++; continue; }
      const bool mine = (batch % n_split) == split;
      o_fb = fb; o_n = (int)min((int64_t)FB, f1 - fb);
      fb += FB; batch++;
      if (mine) return true;
    }
  };
  float px[8], pa = 0.0f, pb = 0.0f, pc = 0.0f;
  auto prefetch = [&](int64_t b0, int n) {
#pragma unroll
    for (int k = 0; k < 8; k++) { const int r = xr + 4 * k; px[k] = (r < n && xc < dim) ? __ldg(feats + (b0 + r) * dim + xc) : 0.0f; }
    const bool ok = af < n && af < FB && d0 + ad < dim;
    pa = ok ? __ldg(ab + (size_t)(b0 + af) * 2 * dim + d0 + ad) : 0.0f;
    pb = ok ? __ldg(ab + (size_t)(b0 + af) * 2 * dim + dim + d0 + ad) : 0.0f;
    pc = t < n ? __ldg(cnt + b0 + t) : 0.0f;
  };
  int64_t cur_fb = 0; int cur_n = 0;
  bool more = next_batch(cur_fb, cur_n);
  if (more) prefetch(cur_fb, cur_n);
  while (more) {
    const int n = cur_n;
    __syncthreads();   // the previous batch has been consumed
#pragma unroll
    for (int k = 0; k < 8; k++) { const int r = xr + 4 * k; if (xc < dim) s_xp[r][xc] = (double)px[k]; else if (xc == dim) s_xp[r][xc] = 1.0; }
    if (af < FB) { s_a[af][ad] = (double)pa; s_b[af][ad] = (double)pb; }
    if (t < FB) s_c[t] = t < n ? pc : 0.0f;
    __syncthreads();
    more = next_batch(cur_fb, cur_n);
    if (more) prefetch(cur_fb, cur_n);
    if (dtile == 0 && t < 32) { const double c = warp_sum_d((double)s_c[t]); if (t == 0) beta += c; }
    // fully unrolled over the batch: every shared-memory operand is (a per-thread base register) + (frame * row stride) as an immediate
    // -- no address arithmetic, no per-frame predicate recomputation in the loop (the first version spent 2/3 of its issue slots on them).
    // Frames beyond n and dropped frames (weight 0: Kaldi never sees them) carry s_c = 0.
#pragma unroll
    for (int f = 0; f < FB; f++) {
      if (s_c[f] != 0.0f) {
        double b[DT];
#pragma unroll
        for (int d = 0; d < DT; d += 2) { const double2 v = *reinterpret_cast<const double2 *>(&s_b[f][d]); b[d] = v.x; b[d + 1] = v.y; }
#pragma unroll
        for (int k = 0; k < NIJ; k++) {
          const double z = xi[k][f * XS] * xj[k][f * XS];
#pragma unroll
          for (int d = 0; d < DT; d++) g[k][d] += b[d] * z;
        }
        kacc[0] += ka0[f * DT] * kx0[f * XS];
        kacc[1] += ka1[f * DT] * kx1[f * XS];
      }
    }
  }
  double *st = stats + (size_t)spk * stats_stride;
  double *K = st + 1, *G = K + (size_t)dim * D1;
#pragma unroll
  for (int k = 0; k < NIJ; k++) {
    const int ij = t + k * ANT;
    if (ij >= NP) continue;
#pragma unroll
    for (int d = 0; d < DT; d++)
      if (d0 + d < dim && g[k][d] != 0.0) atomicAdd(&G[(size_t)(d0 + d) * NP + ij], g[k][d]);
  }
  if (k0ok && d0 + kd0 < dim && kacc[0] != 0.0) atomicAdd(&K[(size_t)(d0 + kd0) * D1 + kj0], kacc[0]);
  if (k1ok && d0 + kd1 < dim && kacc[1] != 0.0) atomicAdd(&K[(size_t)(d0 + kd1) * D1 + kj1], kacc[1]);
  if (t == 0 && dtile == 0 && beta != 0.0) atomicAdd(&st[0], beta);
}

// ---------------------------------------------------------------------------------------------- transform update
// One CTA per speaker: Kaldi ComputeFmllrMatrixDiagGmmFull (transform/fmllr-diag-gmm.cc) in f64 -- invert the D second-order
// matrices G_d (warp-level Gauss-Jordan in shared memory, SPD so unpivoted), then num_iters sweeps of FmllrInnerUpdate over the
// rows: cofactor row = column d of A^-1 (A^-1 is re-inverted with partial pivoting once per sweep and carried across the row
// updates of a sweep with the Sherman-Morrison identity), the quadratic for the step size, w_d = (alpha c + k_d) G_d^-1.  The
// auxiliary function before / after decides whether the estimate is kept (FmllrAuxFuncDiagGmm, compared in float like Kaldi).
constexpr int UNT = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[i] = sum_j M[i][j] v[j] for an n x n row-major matrix in global memory; v, out in shared memory.  Eight lanes share a row (32 rows
// per pass over the CTA's 256 threads), partial sums meet in three shuffles.
__device__ __forceinline__ void matvec_rows(const double *M, const double *v, double *out, int n, int t) {
  const int sub = t & 7;
#pragma unroll
  for (int pass = 0; pass < 3; pass++) {   // n <= 65 < 96
    const int i = (t >> 3) + 32 * pass;
    if (32 * pass >= n) break;             // uniform
    double a = 0.0;
    if (i < n) {
      const double *row = M + (size_t)i * n;   // plain loads: this CTA wrote M earlier in the kernel
      for (int j = sub; j < n; j += 8) a += row[j] * v[j];
    }
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (i < n && sub == 0) out[i] = a;
  }
}

// in-place inverse of the n x n matrix A (shared memory) by Gauss-Jordan with partial pivoting, all threads of the CTA;
// returns log|det A| (valid in every thread).  s_col: n doubles, s_piv: n ints, s_ld: 1 double of scratch.
__device__ double cta_invert(double *A, int n, double *s_col, int *s_piv, double *s_ld) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) *s_ld = 0.0;
  for (int c = 0; c < n; c++) {
    if (warp == 0) {
      double best = -1.0; int bi = c;
      for (int r = c + lane; r < n; r += 32) { const double v = fabs(A[r * n + c]); if (v > best) { best = v; bi = r; } }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) s_piv[c] = bi;
    }
    __syncthreads();
    const int p = s_piv[c];
    if (p != c && t < n) { const double a = A[c * n + t]; A[c * n + t] = A[p * n + t]; A[p * n + t] = a; }
    __syncthreads();
    const double pv = A[c * n + c], ipv = 1.0 / pv;
    __syncthreads();
    if (t == 0) *s_ld += log(fabs(pv));
    if (t < n) { s_col[t] = A[t * n + c]; A[c * n + t] = (t == c) ? ipv : A[c * n + t] * ipv; }
    __syncthreads();
    for (int idx = t; idx < n * n; idx += UNT) {
      const int r = idx / n, j = idx - r * n;
      if (r != c) A[idx] = (j == c) ? -s_col[r] * ipv : A[idx] - s_col[r] * A[c * n + j];
    }
    __syncthreads();
  }
  for (int c = n - 1; c >= 0; c--) {
    const int p = s_piv[c];
    if (p != c && t < n) { const double a = A[t * n + c]; A[t * n + c] = A[t * n + p]; A[t * n + p] = a; }
    __syncthreads();
  }
  return *s_ld;
}

__device__ double cta_sum(double v, double *s_red) {
  const int t = threadIdx.x;
  v = warp_sum(v);
  __syncthreads();
  if ((t & 31) == 0) s_red[t >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < UNT / 32; w++) r += s_red[w];
  return r;
}

// beta log|det A| + tr(W K^T) - 1/2 sum_d w_d G_d w_d^T  (logdet supplied by the caller)
__device__ double cta_auxf(const double *W, const double *K, const double *Gp, double beta, double logdet, int D, double *s_red) {
  const int D1 = D + 1, NP = D1 * (D1 + 1) / 2, t = threadIdx.x;
  double part = 0.0;
  for (int idx = t; idx < D * D1; idx += UNT) part += W[idx] * K[idx];
  for (int idx = t; idx < D * NP; idx += UNT) {
    const int d = idx / NP, ij = idx - d * NP;
    int i = (int)((sqrtf(8.0f * (float)ij + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= ij) i++;
    while (i * (i + 1) / 2 > ij) i--;
    const int j = ij - i * (i + 1) / 2;
    const double q = Gp[idx] * W[d * D1 + i] * W[d * D1 + j];
    part -= (i == j ? 0.5 : 1.0) * q;
  }
  return beta * logdet + cta_sum(part, s_red);
}

__global__ void __launch_bounds__(UNT)
fmllr_update_kernel(const double *__restrict__ stats, int64_t stats_stride, int dim, int num_iters, double min_count, int n_par,
                    double *invG_all, float *__restrict__ W_out, double *__restrict__ impr_out, double *__restrict__ count_out) {
  extern __shared__ double sm[];
  const int D = dim, D1 = D + 1, NP = D1 * (D1 + 1) / 2;
  const int spk = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const double *st = stats + (size_t)spk * stats_stride;
  const double beta = st[0], *K = st + 1, *Gp = K + (size_t)D * D1;
  float *Wo = W_out + (size_t)spk * D * D1;
  if (t == 0) { count_out[spk] = beta; impr_out[spk] = 0.0; }
  for (int idx = t; idx < D * D1; idx += UNT) Wo[idx] = (idx / D1 == idx % D1) ? 1.0f : 0.0f;
  if (!(beta > min_count)) return;   // gmm-est-fmllr: below min-count the unit transform is written
  double *invG = invG_all + (size_t)spk * D * D1 * D1;
  // ---- phase A: G_d^-1, one matrix per warp at a time
  if (warp < n_par) {
    double *M = sm + (size_t)warp * D1 * D1;
    for (int d = warp; d < D; d += n_par) {
      const double *g = Gp + (size_t)d * NP;
      for (int idx = lane; idx < D1 * D1; idx += 32) {
        const int i = idx / D1, j = idx - i * D1;
        M[idx] = j <= i ? g[i * (i + 1) / 2 + j] : g[j * (j + 1) / 2 + i];
      }
      __syncwarp();
      for (int c = 0; c < D1; c++) {
        const double ip = 1.0 / M[c * D1 + c];
        __syncwarp();
        for (int j = lane; j < D1; j += 32) M[c * D1 + j] = (j == c) ? ip : M[c * D1 + j] * ip;
        __syncwarp();
        for (int r = 0; r < D1; r++) {
          if (r == c) continue;
          const double f = M[r * D1 + c];
          __syncwarp();
          for (int j = lane; j < D1; j += 32) M[r * D1 + j] = (j == c) ? -f * ip : M[r * D1 + j] - f * M[c * D1 + j];
        }
        __syncwarp();
      }
      double *o = invG + (size_t)d * D1 * D1;
      for (int idx = lane; idx < D1 * D1; idx += 32) o[idx] = M[idx];
      __syncwarp();
    }
  }
  __threadfence_block();
  __syncthreads();
  // ---- phase B: row updates.  Shared memory is re-carved.
  double *W = sm, *Ainv = W + D * D1, *cvec = Ainv + D * D, *cg = cvec + D1, *vvec = cg + D1, *wnew = vvec + D1, *delta = wnew + D1,
         *u = delta + D1, *vv = u + D1, *s_col = vv + D1, *s_red = s_col + D1, *s_ld = s_red + 8;
  int *s_piv = reinterpret_cast<int *>(s_ld + 1);
  for (int idx = t; idx < D * D1; idx += UNT) W[idx] = (idx / D1 == idx % D1) ? 1.0 : 0.0;
  __syncthreads();
  constexpr int MAXE = (64 * 64 + UNT - 1) / UNT;   // D x D elements per thread, D <= 64
  unsigned char ei[MAXE], ej[MAXE];
#pragma unroll
  for (int k = 0; k < MAXE; k++) { const int idx = t + k * UNT; ei[k] = (unsigned char)(idx < D * D ? idx / D : 0); ej[k] = (unsigned char)(idx < D * D ? idx % D : 0); }
  // objective at the unit transform: log|det| = 0
  const double old_objf = (double)(float)cta_auxf(W, K, Gp, beta, 0.0, D, s_red);
  for (int it = 0; it < num_iters; it++) {
    for (int idx = t; idx < D * D; idx += UNT) { const int r = idx / D, j = idx - r * D; Ainv[idx] = W[r * D1 + j]; }
    __syncthreads();
    cta_invert(Ainv, D, s_col, s_piv, s_ld);
    for (int d = 0; d < D; d++) {
      const double *iG = invG + (size_t)d * D1 * D1, *k = K + (size_t)d * D1;
      if (t < D1) cvec[t] = t < D ? Ainv[t * D + d] : 0.0;
      __syncthreads();
      matvec_rows(iG, cvec, cg, D1, t);
      __syncthreads();
      double p1 = 0.0, p2 = 0.0;
      for (int j = lane; j < D1; j += 32) { p1 += cg[j] * cvec[j]; p2 += cg[j] * k[j]; }
      const double e1 = warp_sum(p1), e2 = warp_sum(p2);
      const double disc = sqrt(e2 * e2 + 4.0 * e1 * beta);
      const double a1 = (-e2 + disc) / (2.0 * e1), a2 = (-e2 - disc) / (2.0 * e1);
      const double f1 = beta * log(fabs(a1 * e1 + e2)) - 0.5 * a1 * a1 * e1, f2 = beta * log(fabs(a2 * e1 + e2)) - 0.5 * a2 * a2 * e1;
      const double alpha = f1 > f2 ? a1 : a2;
      if (t < D1) vvec[t] = alpha * cvec[t] + k[t];
      __syncthreads();
      matvec_rows(iG, vvec, wnew, D1, t);
      __syncthreads();
      if (t < D) { delta[t] = wnew[t] - W[d * D1 + t]; u[t] = Ainv[t * D + d]; }
      __syncthreads();
      if (t < D1) W[d * D1 + t] = wnew[t];
      if (t < D) { double a = 0.0; for (int i = 0; i < D; i++) a += delta[i] * Ainv[i * D + t]; vv[t] = a; }
      __syncthreads();
      const double inv_den = 1.0 / (1.0 + vv[d]);
#pragma unroll
      for (int k = 0; k < MAXE; k++) { const int idx = t + k * UNT; if (idx < D * D) Ainv[idx] -= u[ei[k]] * vv[ej[k]] * inv_den; }
      __syncthreads();
    }
  }
  for (int idx = t; idx < D * D; idx += UNT) { const int r = idx / D, j = idx - r * D; Ainv[idx] = W[r * D1 + j]; }
  __syncthreads();
  const double logdet = cta_invert(Ainv, D, s_col, s_piv, s_ld);
  const double new_objf = (double)(float)cta_auxf(W, K, Gp, beta, logdet, D, s_red);
  const double impr = new_objf - old_objf;
  const bool approx_equal = fabs(new_objf - old_objf) <= 0.001 * (fabs(new_objf) + fabs(old_objf));
  if (impr < 0.0 && !approx_equal) return;   // "objective function did not increase": keep the unit transform
  for (int idx = t; idx < D * D1; idx += UNT) Wo[idx] = (float)W[idx];
  if (t == 0) impr_out[spk] = impr;
}
}  // namespace

extern "C" int64_t mfa_fmllr_stats_size(int32_t dim) {
  const int64_t D1 = dim + 1;
  return 1 + (int64_t)dim * D1 + (int64_t)dim * (D1 * (D1 + 1) / 2);
}

namespace mfa {
int launch_fmllr_acc(mfa_engine *e, mfa_model *mp, mfa_model *ms, const float *d_feats, const int32_t *d_ali, const float *d_tid_weight,
                     const int64_t *d_frame_off, const int64_t *h_frame_off, const int32_t *h_utt2spk, int32_t n_utts, int32_t n_spk,
                     double *d_stats) {
  const int dim = ms->dim;
  if (mp->dim != dim || mp->num_gauss != ms->num_gauss || mp->num_pdfs != ms->num_pdfs || mp->h_pdf_off != ms->h_pdf_off)
    return set_error(MFA_ERR_INVALID, "fMLLR: the posterior model and the statistics model must share one Gaussian layout");
  // the accumulation kernel stages [x | 1] in 64-wide rows: the bias column needs a free slot, so dim 64 is out
  if (dim > 63) return set_error(MFA_ERR_UNSUPPORTED, "fMLLR: dim > 63");
  for (int p = 0; p < ms->num_pdfs; p++)
    if (ms->h_pdf_off[p + 1] - ms->h_pdf_off[p] > MFA_TILE_N) return set_error(MFA_ERR_UNSUPPORTED, "fMLLR: pdf with more than 128 components");
  const int64_t n_frames = h_frame_off[n_utts];
  if (n_frames == 0 || n_spk == 0) return MFA_OK;
  std::vector<int32_t> off(n_spk + 1, 0), utts(n_utts);
  for (int u = 0; u < n_utts; u++) { const int s = h_utt2spk[u]; if (s < 0 || s >= n_spk) return set_error(MFA_ERR_INVALID, "utt2spk out of range"); off[s + 1]++; }
  for (int s = 0; s < n_spk; s++) off[s + 1] += off[s];
  { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int u = 0; u < n_utts; u++) utts[cur[h_utt2spk[u]]++] = u; }
  int32_t *d_off, *d_utts; float *d_ab;
  MFA_TRY(e->upload(DB_SPK_UTT_OFF, off.data(), off.size(), &d_off));
  MFA_TRY(e->upload(DB_SPK_UTTS, utts.data(), utts.size(), &d_utts));
  MFA_TRY(e->getT<float>(DB_FM_AUX, (size_t)n_frames * (2 * dim + 1), &d_ab));
  float *d_cnt = d_ab + (size_t)n_frames * 2 * dim;
  int64_t blocks = (n_frames + FW - 1) / FW;
  blocks = std::min<int64_t>(blocks, (int64_t)e->sm_count * 16);
  fmllr_frame_kernel<<<(unsigned)blocks, FW * 32, 0, e->stream>>>(d_feats, d_ali, n_frames, dim, ms->num_tids, ms->d_pdf_off, ms->d_tid2pdf,
                                                                   d_tid_weight, mp->d_gconsts, mp->d_miv, mp->d_iv, ms->d_miv, ms->d_iv, d_ab, d_cnt);
  const int n_dtile = (dim + DT - 1) / DT;
  int n_split = (int)std::max<int64_t>(1, std::min<int64_t>(64, ((int64_t)e->sm_count * 4 + (int64_t)n_spk * n_dtile - 1) / ((int64_t)n_spk * n_dtile)));
  const int64_t grid = (int64_t)n_spk * n_dtile * n_split;
  const int NP = (dim + 1) * (dim + 2) / 2;
  const int64_t stride = mfa_fmllr_stats_size(dim);
  if (NP <= 4 * ANT)
    fmllr_accum_kernel<4><<<(unsigned)grid, ANT, 0, e->stream>>>(d_feats, d_ab, d_cnt, dim, d_frame_off, d_off, d_utts, n_dtile, n_split, d_stats, stride);
  else
    fmllr_accum_kernel<9><<<(unsigned)grid, ANT, 0, e->stream>>>(d_feats, d_ab, d_cnt, dim, d_frame_off, d_off, d_utts, n_dtile, n_split, d_stats, stride);
  e->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

int launch_fmllr_update(mfa_engine *e, const double *d_stats, int dim, int32_t n_spk, int num_iters, double min_count, float *d_W,
                        double *d_impr, double *d_count) {
  if (n_spk == 0) return MFA_OK;
  if (dim > 64 || dim < 1) return set_error(MFA_ERR_UNSUPPORTED, "fMLLR: dim must be in 1..64");
  const size_t D1 = dim + 1;
  double *d_invG;
  MFA_TRY(e->getT<double>(DB_FM_INVG, (size_t)n_spk * dim * D1 * D1, &d_invG));
  const size_t phase_b = sizeof(double) * ((size_t)dim * D1 + (size_t)dim * dim + 9 * D1 + 16) + sizeof(int) * D1;
  int n_par = 8;
  while (n_par > 1 && sizeof(double) * n_par * D1 * D1 > e->smem_optin - 2048) n_par--;
  const size_t smem = std::max(phase_b, sizeof(double) * n_par * D1 * D1);
  CUDA_TRY(cudaFuncSetAttribute(fmllr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fmllr_update_kernel<<<(unsigned)n_spk, UNT, smem, e->stream>>>(d_stats, mfa_fmllr_stats_size(dim), dim, num_iters, min_count, n_par, d_invG,
                                                                  d_W, d_impr, d_count);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}
}  // namespace mfa
