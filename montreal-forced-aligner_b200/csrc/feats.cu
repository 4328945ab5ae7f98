// feats.cu -- a5: CMVN apply -> (delta+delta-delta | splice + LDA) -> per-speaker fMLLR, one fused kernel.
//
// Replaces kalpy FeatureArchive(cmvn, deltas | splices + lda_mat, transform) built by
// Job.construct_feature_archive (reference: montreal_forced_aligner/db.py:2101-2136; op order restated in-tree at
// alignment/multiprocessing.py:1287-1304).  Semantics per SURVEY.md A.3 (Kaldi transform/cmvn.cc ApplyCmvn with
// norm_vars=false, feat/feature-functions.cc DeltaFeatures / SpliceFrames, transform-common ApplyAffineTransform).
#include <algorithm>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {
constexpr int TF = 32;       // frames per block
constexpr int MAXH = 8;      // max halo (delta order-2 window-2 needs 4)

struct FeatParams {
  int mode, in_dim, ctx, lda_rows, lda_cols, mid_dim, out_dim, has_fmllr, has_cmvn, halo;
  float s1[5], s2[9];
};

__global__ void __launch_bounds__(128)
feat_kernel(FeatParams p, const float *__restrict__ in, const int64_t *__restrict__ frame_off, const int64_t *__restrict__ row_off,
            const int32_t *__restrict__ utt2spk, const int64_t *__restrict__ tile_off, int n_utts, const float *__restrict__ lda,
            const float *__restrict__ fmllr, const double *__restrict__ cmvn, float *__restrict__ out, int out_ld) {
  extern __shared__ float sm[];
  float *raw = sm;                                          // [(TF+2*halo)][in_dim]
  float *mid = raw + (TF + 2 * p.halo) * p.in_dim;          // [TF][mid_dim]
  float *mat = mid + TF * p.mid_dim;                        // lda or fmllr staging
  const int64_t b = blockIdx.x;
  int lo = 0, hi = n_utts - 1;
  while (lo < hi) { int m = (lo + hi + 1) >> 1; if (tile_off[m] <= b) lo = m; else hi = m - 1; }
  const int u = lo;
  const int64_t f0 = frame_off[u], T = frame_off[u + 1] - f0;
  const int64_t t0 = (b - tile_off[u]) * TF;
  const int nt = (int)min((int64_t)TF, T - t0);
  const int spk = utt2spk ? utt2spk[u] : 0;
  const int D = p.in_dim;
  // stage raw rows t0-halo .. t0+TF+halo (clamped to the utterance), CMVN offset applied
  for (int i = threadIdx.x; i < (TF + 2 * p.halo) * D; i += blockDim.x) {
    int r = i / D, d = i % D;
    int64_t t = t0 + r - p.halo;
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    float v = in[(f0 + t) * D + d];
    if (p.has_cmvn) {
      const double *st = cmvn + (size_t)spk * 2 * (D + 1);
      float off = (float)(-(st[d] / st[D]));
      v += off;
    }
    raw[i] = v;
  }
  if (p.mode == 2) for (int i = threadIdx.x; i < p.lda_rows * p.lda_cols; i += blockDim.x) mat[i] = lda[i];
  __syncthreads();
  if (p.mode == 0) {
    for (int i = threadIdx.x; i < nt * D; i += blockDim.x) mid[(i / D) * p.mid_dim + i % D] = raw[(i / D + p.halo) * D + i % D];
  } else if (p.mode == 1) {
    for (int i = threadIdx.x; i < nt * D; i += blockDim.x) {
      int r = i / D, d = i % D;
      const float *c = raw + (r + p.halo) * D + d;
      float d1 = 0.0f, d2 = 0.0f;
#pragma unroll
      for (int j = -2; j <= 2; j++) if (p.s1[j + 2] != 0.0f) d1 += p.s1[j + 2] * c[j * D];
#pragma unroll
      for (int j = -4; j <= 4; j++) if (p.s2[j + 4] != 0.0f) d2 += p.s2[j + 4] * c[j * D];
      float *o = mid + r * p.mid_dim;
      o[d] = c[0]; o[D + d] = d1; o[2 * D + d] = d2;
    }
  } else {
    const int w = 2 * p.ctx + 1, sd = w * D;
    for (int i = threadIdx.x; i < nt * p.lda_rows; i += blockDim.x) {
      int r = i / p.lda_rows, k = i % p.lda_rows;
      const float *x = raw + (r + p.halo - p.ctx) * D;  // spliced vector is contiguous rows r-ctx..r+ctx
      const float *m = mat + k * p.lda_cols;
      float acc = 0.0f;
      for (int c = 0; c < sd; c++) acc += m[c] * x[c];
      if (p.lda_cols == sd + 1) acc += m[sd];
      mid[r * p.mid_dim + k] = acc;
    }
  }
  __syncthreads();
  const int64_t orow = row_off[u] + t0;
  if (p.has_fmllr) {
    const int Dm = p.mid_dim;
    const float *A = fmllr + (size_t)spk * Dm * (Dm + 1);
    for (int i = threadIdx.x; i < Dm * (Dm + 1); i += blockDim.x) mat[i] = A[i];
    __syncthreads();
    for (int i = threadIdx.x; i < nt * Dm; i += blockDim.x) {
      int r = i / Dm, k = i % Dm;
      const float *m = mat + k * (Dm + 1), *x = mid + r * Dm;
      float acc = 0.0f;
      for (int c = 0; c < Dm; c++) acc += m[c] * x[c];
      acc += m[Dm];
      out[(orow + r) * out_ld + k] = acc;
    }
  } else {
    for (int i = threadIdx.x; i < nt * p.mid_dim; i += blockDim.x) out[(orow + i / p.mid_dim) * out_ld + i % p.mid_dim] = mid[i];
  }
}

// ---- splice + LDA (+ fMLLR) as a small register-tiled GEMM ------------------------------------------------------------------
// Persistent CTAs: the LDA matrix is staged once per CTA, transposed to [column][row] so that a thread reads four consecutive
// output rows with one 16-byte load; every thread owns a 4-frame x 4-output register tile (per spliced column: 4 scalar reads of
// the staged frames + 1 vector read of the matrix feed 16 FMAs).  Tiles never cross an utterance.
constexpr int LT = 256;      // threads per CTA

__global__ void __launch_bounds__(LT)
feat_lda_kernel(FeatParams p, int tf, int n_kq, const float *__restrict__ in, const int64_t *__restrict__ frame_off, const int64_t *__restrict__ row_off,
                const int32_t *__restrict__ utt2spk, const int64_t *__restrict__ tile_off, int n_utts, int64_t n_tiles, const float *__restrict__ lda,
                const float *__restrict__ fmllr, const double *__restrict__ cmvn, float *__restrict__ out, int out_ld) {
  extern __shared__ __align__(16) float sm[];
  const int D = p.in_dim, sd = (2 * p.ctx + 1) * D, KP = 4 * n_kq;          // KP: output rows padded to a multiple of 4
  float *mt = sm;                                                            // [sd + 1][KP]  transposed LDA (+ offset row)
  float *raw = mt + (size_t)(sd + 1) * KP;                                   // [tf + 2 ctx][D]
  float *mid = raw + (((size_t)(tf + 2 * p.ctx) * D + 3) & ~(size_t)3);      // [tf][KP] (fMLLR only), 16-byte aligned
  float *fm = mid + (p.has_fmllr ? (size_t)tf * KP : 0);                     // [mid_dim][mid_dim + 1] (fMLLR only)
  for (int i = threadIdx.x; i < (sd + 1) * KP; i += LT) {
    const int c = i / KP, k = i % KP;
    mt[i] = (k < p.lda_rows && c < p.lda_cols) ? lda[(size_t)k * p.lda_cols + c] : 0.0f;
  }
  const int kq = threadIdx.x % n_kq, fq = threadIdx.x / n_kq;               // output quad, frame quad
  const bool active = fq * 4 < tf;
  int u = 0;
  for (int64_t b = blockIdx.x; b < n_tiles; b += gridDim.x) {
    { int lo = u, hi = n_utts - 1; while (lo < hi) { int m = (lo + hi + 1) >> 1; if (tile_off[m] <= b) lo = m; else hi = m - 1; } u = lo; }
    const int64_t f0 = frame_off[u], T = frame_off[u + 1] - f0;
    const int64_t t0 = (b - tile_off[u]) * tf;
    const int nt = (int)min((int64_t)tf, T - t0);
    const int spk = utt2spk ? utt2spk[u] : 0;
    __syncthreads();                                                         // previous tile's readers are done (first pass: mt staged)
    for (int i = threadIdx.x; i < (tf + 2 * p.ctx) * D; i += LT) {
      const int r = i / D, d = i % D;
      int64_t t = t0 + r - p.ctx;
      t = t < 0 ? 0 : (t >= T ? T - 1 : t);
      float v = in[(f0 + t) * D + d];
      if (p.has_cmvn) { const double *st = cmvn + (size_t)spk * 2 * (D + 1); v += (float)(-(st[d] / st[D])); }
      raw[i] = v;
    }
    if (p.has_fmllr) {
      const int Dm = p.mid_dim;
      const float *A = fmllr + (size_t)spk * Dm * (Dm + 1);
      for (int i = threadIdx.x; i < Dm * (Dm + 1); i += LT) fm[i] = A[i];
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int k = 0; k < 4; k++) acc[f][k] = 0.0f;
    if (active) {
      const float *x0 = raw + (size_t)(fq * 4) * D;                          // spliced vector of frame r = staged rows r .. r + 2 ctx, contiguous
      const float4 *m4 = (const float4 *)(mt + 4 * kq);
      for (int c = 0; c < sd; c++) {
        const float4 m = m4[(size_t)c * n_kq];
        const float xa = x0[c], xb = x0[c + D], xc = x0[c + 2 * D], xd = x0[c + 3 * D];
        acc[0][0] += m.x * xa; acc[0][1] += m.y * xa; acc[0][2] += m.z * xa; acc[0][3] += m.w * xa;
        acc[1][0] += m.x * xb; acc[1][1] += m.y * xb; acc[1][2] += m.z * xb; acc[1][3] += m.w * xb;
        acc[2][0] += m.x * xc; acc[2][1] += m.y * xc; acc[2][2] += m.z * xc; acc[2][3] += m.w * xc;
        acc[3][0] += m.x * xd; acc[3][1] += m.y * xd; acc[3][2] += m.z * xd; acc[3][3] += m.w * xd;
      }
      if (p.lda_cols == sd + 1) {
        const float4 m = m4[(size_t)sd * n_kq];
#pragma unroll
        for (int f = 0; f < 4; f++) { acc[f][0] += m.x; acc[f][1] += m.y; acc[f][2] += m.z; acc[f][3] += m.w; }
      }
    }
    const int64_t orow = row_off[u] + t0;
    if (!p.has_fmllr) {
      if (active) {
#pragma unroll
        for (int f = 0; f < 4; f++) {
          const int r = fq * 4 + f;
          if (r < nt) {
            float *o = out + (orow + r) * out_ld + 4 * kq;
            if (4 * kq + 3 < p.lda_rows && (out_ld & 3) == 0) *(float4 *)o = make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
            else for (int k = 0; k < 4; k++) if (4 * kq + k < p.lda_rows) o[k] = acc[f][k];
          }
        }
      }
    } else {
      if (active) {
#pragma unroll
        for (int f = 0; f < 4; f++) *(float4 *)(mid + (size_t)(fq * 4 + f) * KP + 4 * kq) = make_float4(acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
      }
      __syncthreads();
      const int Dm = p.mid_dim;
      for (int i = threadIdx.x; i < nt * Dm; i += LT) {
        const int r = i / Dm, k = i % Dm;
        const float *m = fm + k * (Dm + 1), *x = mid + (size_t)r * KP;
        float a = 0.0f;
        for (int c = 0; c < Dm; c++) a += m[c] * x[c];
        out[(orow + r) * out_ld + k] = a + m[Dm];
      }
    }
  }
}

__global__ void transpose_kernel(const float *__restrict__ in, int64_t rows, int64_t cols, int64_t in_ld, float *__restrict__ out, int64_t out_ld) {
  __shared__ float tile[32][33];
  int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * in_ld + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * out_ld + r] = tile[threadIdx.x][i];
  }
}
}  // namespace

extern "C" int32_t mfa_feat_out_dim(const mfa_feat_opts *o) {
  if (o->mode == 0) return o->in_dim;
  if (o->mode == 1) return 3 * o->in_dim;
  return o->lda_rows;
}

namespace mfa {

int launch_features(mfa_engine *e, const mfa_feat_opts *o, const float *d_in, const int64_t *d_frame_off, const int64_t *h_frame_off,
                    const int64_t *d_row_off, const int32_t *d_utt2spk, int32_t n_utts, const double *d_cmvn_stats, float *d_out, int out_ld) {
  if (n_utts == 0) return MFA_OK;
  FeatParams p{};
  p.mode = o->mode; p.in_dim = o->in_dim; p.ctx = o->splice_ctx; p.lda_rows = o->lda_rows; p.lda_cols = o->lda_cols;
  p.has_fmllr = o->fmllr != nullptr; p.has_cmvn = d_cmvn_stats != nullptr;
  if (p.in_dim < 1 || p.in_dim > 64) return set_error(MFA_ERR_UNSUPPORTED, "in_dim must be in 1..64");
  if (p.mode == 0) { p.halo = 0; p.mid_dim = p.in_dim; }
  else if (p.mode == 1) { p.halo = 4; p.mid_dim = 3 * p.in_dim; }
  else if (p.mode == 2) {
    if (p.ctx < 0 || p.ctx > MAXH) return set_error(MFA_ERR_UNSUPPORTED, "splice context must be in 0..8");
    int sd = (2 * p.ctx + 1) * p.in_dim;
    if (!o->lda || (p.lda_cols != sd && p.lda_cols != sd + 1) || p.lda_rows < 1 || p.lda_rows > 256) return set_error(MFA_ERR_INVALID, "LDA matrix shape does not match spliced dim");
    p.halo = p.ctx; p.mid_dim = p.lda_rows;
  } else return set_error(MFA_ERR_INVALID, "bad feature mode");
  p.out_dim = p.mid_dim;
  const float s1[5] = {-0.2f, -0.1f, 0.0f, 0.1f, 0.2f};
  for (int i = 0; i < 5; i++) p.s1[i] = s1[i];
  for (int i = 0; i < 9; i++) p.s2[i] = 0.0f;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) p.s2[i + j] += s1[i] * s1[j];
  if (p.has_fmllr && o->n_spk < 1) return set_error(MFA_ERR_INVALID, "fmllr given but n_spk < 1");
  // splice + LDA: register-tiled persistent kernel when the geometry fits (any LDA with <= 256 rows does)
  const int n_kq = p.mode == 2 ? (p.lda_rows + 3) / 4 : 0;
  const bool tiled = p.mode == 2 && n_kq >= 1 && n_kq <= 64;
  const int tf = tiled ? 4 * (LT / n_kq) : TF;                     // frames per tile
  // tile prefix
  std::vector<int64_t> tile_off(n_utts + 1, 0);
  for (int u = 0; u < n_utts; u++) tile_off[u + 1] = tile_off[u] + (h_frame_off[u + 1] - h_frame_off[u] + tf - 1) / tf;
  int64_t n_tiles = tile_off[n_utts];
  if (n_tiles == 0) return MFA_OK;
  int64_t *d_tile_off;
  MFA_TRY(e->upload(DB_TILE_OFF, tile_off.data(), tile_off.size(), &d_tile_off));
  float *d_lda = nullptr, *d_fmllr = nullptr;
  if (p.mode == 2) MFA_TRY(e->upload(DB_LDA, o->lda, (size_t)p.lda_rows * p.lda_cols, &d_lda));
  if (p.has_fmllr) MFA_TRY(e->upload(DB_FMLLR, o->fmllr, (size_t)o->n_spk * p.mid_dim * (p.mid_dim + 1), &d_fmllr));
  if (tiled) {
    const int sd = (2 * p.ctx + 1) * p.in_dim, KP = 4 * n_kq;
    const size_t smem = ((size_t)(sd + 1) * KP + (((size_t)(tf + 2 * p.ctx) * p.in_dim + 3) & ~(size_t)3) + (p.has_fmllr ? (size_t)tf * KP + (size_t)p.mid_dim * (p.mid_dim + 1) : 0)) * sizeof(float);
    if (smem <= e->smem_optin) {
      CUDA_TRY(cudaFuncSetAttribute(feat_lda_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (e->smem_optin) / (smem + 1024)));
      const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)e->sm_count * per_sm);
      feat_lda_kernel<<<grid, LT, smem, e->stream>>>(p, tf, n_kq, d_in, d_frame_off, d_row_off, d_utt2spk, d_tile_off, n_utts, n_tiles, d_lda, d_fmllr,
                                                     d_cmvn_stats, d_out, out_ld);
      e->launches++;
      CUDA_TRY(cudaGetLastError());
      return MFA_OK;
    }
    return set_error(MFA_ERR_UNSUPPORTED, "splice + LDA tile exceeds shared memory");
  }
  size_t mat = 0;
  if (p.mode == 2) mat = (size_t)p.lda_rows * p.lda_cols;
  if (p.has_fmllr) mat = std::max(mat, (size_t)p.mid_dim * (p.mid_dim + 1));
  size_t smem = ((size_t)(TF + 2 * p.halo) * p.in_dim + (size_t)TF * p.mid_dim + mat) * sizeof(float);
  if (smem > e->smem_optin) return set_error(MFA_ERR_UNSUPPORTED, "feature kernel shared memory exceeds the device limit");
  CUDA_TRY(cudaFuncSetAttribute(feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  feat_kernel<<<(unsigned)n_tiles, 128, smem, e->stream>>>(p, d_in, d_frame_off, d_row_off, d_utt2spk, d_tile_off, n_utts, d_lda, d_fmllr,
                                                           d_cmvn_stats, d_out, out_ld);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

int launch_transpose(mfa_engine *e, const float *d_in, int64_t rows, int64_t cols, int64_t in_ld, float *d_out, int64_t out_ld) {
  if (rows == 0 || cols == 0) return MFA_OK;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  if (grid.y > 65535) return set_error(MFA_ERR_UNSUPPORTED, "transpose: too many rows for one launch");
  transpose_kernel<<<grid, dim3(32, 8), 0, e->stream>>>(d_in, rows, cols, in_ld, d_out, out_ld);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

}  // namespace mfa
