// accstats.cu -- K4: GMM accumulator statistics for the align -> acc-stats -> update training loop.
//
// Replaces GmmStatsAccumulator.accumulate_stats / AccumAmDiagGmm.acc_stats + TransitionModel.acc_stats
// (reference call sites: montreal_forced_aligner/alignment/multiprocessing.py:652-666,
// acoustic_modeling/monophone.py:114-120).  Semantics per SURVEY.md A.8 (Kaldi gmmbin/gmm-acc-stats-ali.cc,
// gmm/mle-diag-gmm.cc AccumulateFromPosteriors, gmm/diag-gmm.cc ComponentPosteriors): fp32 component
// log-likelihoods and soft-max posteriors of the ALIGNED pdf only, f64 accumulators
//   occ[m] += g, mean_acc[m,:] += g x, var_acc[m,:] += g x^2, trans_acc[tid] += 1, tot_like += logsumexp.
//
// Primary path ("segmented"): frames are counting-sorted by aligned pdf on the device (histogram, one-CTA scan, scatter),
// cut into items of <= ACC_F frames of ONE pdf; a persistent grid walks contiguous item ranges.  Per item the CTA stages the
// pdf's Gaussians (reused while consecutive items share the pdf) and the item's feature rows in shared memory, computes the
// posteriors lane-per-component, and every thread owns (component, dimension) pairs whose f64 sums over the item's frames
// live in registers; one f64 red per (pair, item) reaches HBM instead of one per (pair, frame): T/ACC_F + P flushes of
// nm*(2D+1) doubles instead of T.  Transition counts are an exact int32 histogram added to the f64 block at the end.
// The first version (one warp per frame, a red.global.add.f64 per component x dimension x frame) is kept behind
// MFA_ACC_IMPL=atomic for cross-checks.
#include <algorithm>
#include <cstdlib>

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

// ---------------------------------------------------------------------------------------------- segmented path
constexpr int ACC_F = 128;   // frames per item
constexpr int ACC_NT = 256;  // threads per CTA

// ints[]: pdf_count[P] | tid_count[num_tids+1] | pdf_start[P+1] | cursor[P] | item_off[P+1] | n_items
__global__ void acc_hist_kernel(const int32_t *__restrict__ ali, int64_t n_frames, int num_tids, const int32_t *__restrict__ tid2pdf,
                                int32_t *__restrict__ pdf_count, int32_t *__restrict__ tid_count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += stride) {
    const int tid = ali[f];
    if (tid <= 0 || tid > num_tids) continue;
    const int pdf = tid2pdf[tid];
    // warp-aggregate equal pdfs (silence pdfs take a fifth of all frames)
    const unsigned peers = __match_any_sync(__activemask(), pdf);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&pdf_count[pdf], __popc(peers));
    atomicAdd(&tid_count[tid], 1);
  }
}

__global__ void __launch_bounds__(1024)
acc_scan_kernel(int num_pdfs, int num_tids, const int32_t *__restrict__ pdf_count, const int32_t *__restrict__ tid_count,
                int32_t *__restrict__ pdf_start, int32_t *__restrict__ cursor, int32_t *__restrict__ item_off, int32_t *__restrict__ n_items,
                double *__restrict__ trans) {
  __shared__ int s_cnt[1024], s_itm[1024];
  const int t = threadIdx.x;
  const int per = (num_pdfs + 1023) / 1024;
  const int p0 = t * per, p1 = min(num_pdfs, p0 + per);
  int c = 0, it = 0;
  for (int p = p0; p < p1; p++) { const int n = pdf_count[p]; c += n; it += (n + ACC_F - 1) / ACC_F; }
  s_cnt[t] = c; s_itm[t] = it;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int a = t >= o ? s_cnt[t - o] : 0, b = t >= o ? s_itm[t - o] : 0;
    __syncthreads();
    s_cnt[t] += a; s_itm[t] += b;
    __syncthreads();
  }
  int base_c = s_cnt[t] - c, base_i = s_itm[t] - it;
  for (int p = p0; p < p1; p++) {
    const int n = pdf_count[p];
    pdf_start[p] = base_c; cursor[p] = base_c; item_off[p] = base_i;
    base_c += n; base_i += (n + ACC_F - 1) / ACC_F;
  }
  if (t == 1023) { pdf_start[num_pdfs] = s_cnt[1023]; item_off[num_pdfs] = s_itm[1023]; *n_items = s_itm[1023]; }
  for (int i = t; i <= num_tids; i += 1024) { const int n = tid_count[i]; if (n) trans[i] += (double)n; }
}

__global__ void acc_scatter_kernel(const int32_t *__restrict__ ali, int64_t n_frames, int num_tids, const int32_t *__restrict__ tid2pdf,
                                   int32_t *__restrict__ cursor, int32_t *__restrict__ order) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += stride) {
    const int tid = ali[f];
    if (tid <= 0 || tid > num_tids) continue;
    const int pdf = tid2pdf[tid];
    const unsigned peers = __match_any_sync(__activemask(), pdf);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cursor[pdf], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)f;
  }
}

// dynamic shared memory: g[nmp] | miv[nmp][DS] | iv[nmp][DS] | x[ACC_F][DS] | post[ACC_F][nmp]   (DS = dim | 1: odd stride)
__global__ void __launch_bounds__(ACC_NT)
acc_items_kernel(const float *__restrict__ feats, int dim, int num_pdfs, int num_gauss, int num_tids, int nmp,
                 const int32_t *__restrict__ pdf_off, const float *__restrict__ gconsts, const float *__restrict__ miv,
                 const float *__restrict__ iv, const int32_t *__restrict__ pdf_count, const int32_t *__restrict__ pdf_start,
                 const int32_t *__restrict__ item_off, const int32_t *__restrict__ n_items_p, const int32_t *__restrict__ order,
                 double *__restrict__ acc) {
  extern __shared__ float sm[];
  const int DS = dim | 1;
  float *s_g = sm, *s_miv = s_g + nmp, *s_iv = s_miv + (size_t)nmp * DS, *s_x = s_iv + (size_t)nmp * DS, *s_post = s_x + (size_t)ACC_F * DS;
  __shared__ double s_like[ACC_NT / 32];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  double *occ = acc, *mean = acc + num_gauss, *var = mean + (size_t)num_gauss * dim, *trans = var + (size_t)num_gauss * dim;
  double *tot = trans + (num_tids + 1);
  const int n_items = *n_items_p;
  const int i0 = (int)((int64_t)n_items * blockIdx.x / gridDim.x), i1 = (int)((int64_t)n_items * (blockIdx.x + 1) / gridDim.x);
  if (i0 >= i1) return;
  // pdf of the first item: last p with item_off[p] <= i0 among pdfs that have items
  int lo = 0, hi = num_pdfs;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (item_off[mid] <= i0) lo = mid; else hi = mid; }
  int pdf = lo, staged = -1;
  double like = 0.0, frames = 0.0;
  for (int it = i0; it < i1; it++) {
    while (item_off[pdf + 1] <= it) pdf++;
    const int k = it - item_off[pdf];
    const int f0 = pdf_start[pdf] + k * ACC_F, n = min(ACC_F, pdf_count[pdf] - k * ACC_F);
    const int m0 = pdf_off[pdf], nm = pdf_off[pdf + 1] - m0;
    __syncthreads();   // previous item's accumulation pass is done with s_x / s_post
    if (staged != pdf) {
      for (int i = t; i < nm; i += ACC_NT) s_g[i] = gconsts[m0 + i];
      for (int i = t; i < nm * dim; i += ACC_NT) {
        const int m = i / dim, d = i - m * dim;
        s_miv[m * DS + d] = miv[(size_t)m0 * dim + i];
        s_iv[m * DS + d] = iv[(size_t)m0 * dim + i];
      }
      staged = pdf;
    }
    for (int i = warp; i < n; i += ACC_NT / 32) {
      const float *row = feats + (size_t)order[f0 + i] * dim;
      for (int d = lane; d < dim; d += 32) s_x[i * DS + d] = __ldg(row + d);
    }
    __syncthreads();
    // posteriors: warp per frame, lane per component
    for (int i = warp; i < n; i += ACC_NT / 32) {
      const float *x = s_x + i * DS;
      float mx = -INFINITY;
      for (int mb = 0; mb < nm; mb += 32) {
        const int m = mb + lane;
        float v = -INFINITY;
        if (m < nm) {
          const float *a = s_miv + m * DS, *b = s_iv + m * DS;
          float d1 = 0.0f, d2 = 0.0f;
          for (int d = 0; d < dim; d++) { const float xv = x[d]; d1 = fmaf(a[d], xv, d1); d2 = fmaf(b[d], xv * xv, d2); }
          v = s_g[m] + d1;
          v = v + (-0.5f) * d2;
          s_post[i * nmp + m] = v;
        }
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      __syncwarp();
      float sum = 0.0f;
      for (int m = lane; m < nm; m += 32) { const float ev = expf(s_post[i * nmp + m] - mx); s_post[i * nmp + m] = ev; sum += ev; }
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.0f / sum;
      for (int m = lane; m < nm; m += 32) s_post[i * nmp + m] *= inv;
      if (lane == 0) { like += (double)(mx + logf(sum)); frames += 1.0; }
    }
    __syncthreads();
    // accumulation: thread owns (component, dimension) pairs; f64 sums over the item's frames stay in registers
    for (int idx = t; idx < nm * dim; idx += ACC_NT) {
      const int m = idx / dim, d = idx - m * dim;
      double am = 0.0, av = 0.0;
      for (int i = 0; i < n; i++) {
        const double g = (double)s_post[i * nmp + m], xv = (double)s_x[i * DS + d];
        am += g * xv;
        av += g * (xv * xv);
      }
      atomicAdd(&mean[(size_t)m0 * dim + idx], am);
      atomicAdd(&var[(size_t)m0 * dim + idx], av);
    }
    for (int m = t; m < nm; m += ACC_NT) {
      double o = 0.0;
      for (int i = 0; i < n; i++) o += (double)s_post[i * nmp + m];
      atomicAdd(&occ[m0 + m], o);
    }
  }
  if (lane == 0) s_like[warp] = like;
  __syncthreads();
  // frames is exact in f64; like: fixed-order sum over the CTA's warps, then one red per CTA
  if (t == 0) {
    double l = 0.0;
    for (int w = 0; w < ACC_NT / 32; w++) l += s_like[w];
    atomicAdd(&tot[0], l);
  }
  if (lane == 0 && frames > 0.0) atomicAdd(&tot[1], frames);
}
}  // namespace

extern "C" int64_t mfa_acc_size(const mfa_model *m) {
  if (!m) return 0;
  return (int64_t)m->num_gauss * (1 + 2 * (int64_t)m->dim) + m->num_tids + 1 + 2;
}

extern "C" int mfa_acc_zero(mfa_engine *e, mfa_model *m) {
  if (!e || !m) return set_error(MFA_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  MFA_TRY(e->join_k3());
  size_t bytes = (size_t)mfa_acc_size(m) * sizeof(double);
  MFA_TRY(m->acc_take(bytes));
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

// host accumulators -> the device block (statistics summed on the host by a caller that keeps kalpy's AccumAmDiagGmm objects, and tests)
extern "C" int mfa_acc_write(mfa_engine *e, mfa_model *m, const double *host_in) {
  if (!e || !m || !host_in) return set_error(MFA_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  const size_t bytes = (size_t)mfa_acc_size(m) * sizeof(double);
  MFA_TRY(m->acc_take(bytes));
  CUDA_TRY(cudaMemcpyAsync(m->d_acc, host_in, bytes, cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return MFA_OK;
}

namespace mfa {
static int launch_acc_stats_atomic(mfa_engine *e, mfa_model *m, const float *d_feats, const int32_t *d_ali, int64_t n_frames) {
  int64_t blocks = (n_frames + AW - 1) / AW;
  int64_t cap = (int64_t)e->sm_count * 16;
  if (blocks > cap) blocks = cap;
  acc_stats_kernel<<<(unsigned)blocks, AW * 32, 0, e->stream>>>(d_feats, d_ali, n_frames, m->dim, m->num_gauss, m->num_tids, m->d_pdf_off,
                                                                 m->d_tid2pdf, m->d_gconsts, m->d_miv, m->d_iv, m->d_acc);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

int launch_acc_stats(mfa_engine *e, mfa_model *m, const float *d_feats, const int32_t *d_ali, int64_t n_frames) {
  if (n_frames == 0) return MFA_OK;
  if (!m->d_acc) MFA_TRY(mfa_acc_zero(e, m));
  int max_nm = 1;
  for (int p = 0; p < m->num_pdfs; p++) max_nm = std::max(max_nm, m->h_pdf_off[p + 1] - m->h_pdf_off[p]);
  const int DS = m->dim | 1;
  const size_t smem = sizeof(float) * ((size_t)max_nm + 2 * (size_t)max_nm * DS + (size_t)ACC_F * DS + (size_t)ACC_F * max_nm);
  if (e->cfg.acc_impl == 1 || smem > e->smem_optin - 1024 || n_frames > (int64_t)0x7fffffff) {
    if (max_nm > MFA_TILE_N || m->dim > 64) return set_error(MFA_ERR_UNSUPPORTED, "acc-stats: pdf with too many components / dim > 64");
    return launch_acc_stats_atomic(e, m, d_feats, d_ali, n_frames);
  }
  const int P = m->num_pdfs, NT = m->num_tids;
  const size_t n_int = (size_t)P + (NT + 1) + (P + 1) + P + (P + 1) + 1;
  int32_t *ints, *order;
  MFA_TRY(e->getT<int32_t>(DB_ACC_INT, n_int, &ints));
  MFA_TRY(e->getT<int32_t>(DB_ACC_ORDER, (size_t)n_frames, &order));
  int32_t *pdf_count = ints, *tid_count = pdf_count + P, *pdf_start = tid_count + (NT + 1), *cursor = pdf_start + (P + 1),
          *item_off = cursor + P, *n_items = item_off + (P + 1);
  CUDA_TRY(cudaMemsetAsync(ints, 0, sizeof(int32_t) * ((size_t)P + NT + 1), e->stream));
  const int64_t want = (n_frames + 255) / 256;
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)e->sm_count * 8);
  double *trans = m->d_acc + (size_t)m->num_gauss * (1 + 2 * (size_t)m->dim);
  acc_hist_kernel<<<grid, 256, 0, e->stream>>>(d_ali, n_frames, NT, m->d_tid2pdf, pdf_count, tid_count);
  acc_scan_kernel<<<1, 1024, 0, e->stream>>>(P, NT, pdf_count, tid_count, pdf_start, cursor, item_off, n_items, trans);
  acc_scatter_kernel<<<grid, 256, 0, e->stream>>>(d_ali, n_frames, NT, m->d_tid2pdf, cursor, order);
  // per device / context attribute: set on every launch (several engines on different devices may live in one process)
  CUDA_TRY(cudaFuncSetAttribute(acc_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (e->smem_optin) / (smem + 2048)));
  // upper bound on the number of items: every pdf may leave one partial item
  const int64_t max_items = n_frames / ACC_F + P;
  const unsigned g2 = (unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)e->sm_count * per_sm, max_items));
  acc_items_kernel<<<g2, ACC_NT, smem, e->stream>>>(d_feats, m->dim, P, m->num_gauss, NT, max_nm, m->d_pdf_off, m->d_gconsts, m->d_miv,
                                                    m->d_iv, pdf_count, pdf_start, item_off, n_items, order, m->d_acc);
  e->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}
}  // namespace mfa
