// mstep.cu -- N3: the M-step of the align -> acc-stats -> all-reduce -> update loop on the device.
//
// Replaces, for a device-resident model and accumulator block, what the reference does on the host after summing the
// jobs' accumulators (upstream montreal_forced_aligner/acoustic_modeling/base.py:319-338 `tm.mle_update(transition_accs)`,
// `am.mle_update(gmm_accs, mixup=current_gaussians, power=power)`; monophone.py:275-296): Kaldi
//   gmm/mle-diag-gmm.cc   MleDiagGmmUpdate      (means, variances with a floor, weights, low-count Gaussians removed,
//                                               the last component of a starved pdf is kept)
//   gmm/mle-am-diag-gmm.cc MleAmDiagGmmUpdate   (all pdfs; objective change summed)
//   gmm/am-diag-gmm.cc    SplitByCount / GetSplitTargets (mix-up: greedy allocation by occupancy^power / components)
//   gmm/diag-gmm.cc       Split                 (heaviest component halved, means perturbed by +-perturb*sqrt(var)*N(0,1))
//   hmm/transition-model.cc MleUpdate           (per transition-state: counts -> probabilities, floored)
// The f64 accumulators produced by K4 (and summed across ranks by the NCCL all-reduce) never leave the GPU: four small
// kernels turn them into the next model in the natural layout (pdf_off | weights | gconsts | means_invvars | inv_vars),
// after which gmm_tc.cu rebuilds the K2 operand images from those device arrays.
//
// GetSplitTargets is a sequential priority-queue loop in Kaldi.  Its result is "the B largest keys occ_j^power/(k+1e-10)
// over all (pdf j, k = 1 .. allowed splits of j)"; here that set is found by a bisection on the key value (one CTA) --
// identical allocations except between exactly equal keys, where Kaldi's heap order is unspecified anyway.
// DiagGmm::Split draws from Kaldi's rand()-based RandGauss; here the draws are a counter-based hash of
// (seed, pdf, component, dimension), so split models agree with Kaldi statistically, not bit for bit (hard part 8).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr int MAXC = MFA_TILE_N;   // components per pdf the engine can hold (K2 tiles never split a pdf)

struct MleScal {   // device scalars, copied to the host with the new pdf offsets
  double objf_impr, count, floored, trans_impr, trans_count, like, frames;
  int32_t removed, split, gauss_after, wmax_bad;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- 1: per pdf (one warp): occupancy, which Gaussians are updated / kept
__global__ void __launch_bounds__(256)
mle_pdf_kernel(int P, const int32_t *__restrict__ pdf_off, const double *__restrict__ occ, double min_occ, double min_w, int remove,
               uint8_t *__restrict__ flags, double *__restrict__ state_occ, int32_t *__restrict__ n_keep) {
  const int p = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  const int o0 = pdf_off[p], n = pdf_off[p + 1] - o0;
  double sum = 0.0;
  for (int m = lane; m < n; m += 32) sum += occ[o0 + m];
  sum = warp_sum_d(sum);
  int kept = 0;
  for (int m = lane; m < n; m += 32) {
    const double o = occ[o0 + m];
    const double prob = sum > 0.0 ? o / sum : 1.0 / n;
    const int upd = o > min_occ && prob > min_w;
    const int keep = remove ? upd : 1;
    flags[o0 + m] = (uint8_t)(keep | (upd << 1));
    kept += keep;
  }
  kept = warp_sum_i(kept);
  __syncwarp();
  if (kept == 0) {   // MleDiagGmmUpdate never removes the only component left: walking the indices in order, the LAST one survives
    if (lane == 0) flags[o0 + n - 1] |= 1;
    kept = 1;
  }
  if (lane == 0) { state_occ[p] = sum; n_keep[p] = kept; }
}

// ---- 2: one CTA: mix-up targets (GetSplitTargets as a key-threshold search), new component counts, new pdf offsets
__device__ __forceinline__ int splits_at_least(double c, int max_splits, double theta) {
  // number of k in [1, max_splits] with c / (k + 1e-10) >= theta   (keys decrease with k)
  if (!(c > 0.0) || max_splits <= 0) return 0;
  double q = c / theta - 1.0e-10;
  int k = q >= (double)max_splits ? max_splits : (q < 0.0 ? 0 : (int)q);
  while (k < max_splits && c / ((double)(k + 1) + 1.0e-10) >= theta) k++;
  while (k > 0 && !(c / ((double)k + 1.0e-10) >= theta)) k--;
  return k;
}

__global__ void __launch_bounds__(1024)
mle_plan_kernel(int P, int G_old, const double *__restrict__ state_occ, const int32_t *__restrict__ n_keep, int mixup, double power,
                double min_count, int32_t *__restrict__ n_new, int32_t *__restrict__ new_off, MleScal *__restrict__ scal) {
  __shared__ long long s_red[32];
  __shared__ long long s_bcast;
  __shared__ int s_scan[1024];
  __shared__ int s_base;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  auto block_sum = [&](long long v) -> long long {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (warp == 0) {
      long long x = lane < (int)(blockDim.x >> 5) ? s_red[lane] : 0;
#pragma unroll
      for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) s_bcast = x;
    }
    __syncthreads();
    return s_bcast;
  };
  // exclusive prefix of v over the chunk's threads plus the running base; the base advances by the chunk total
  auto chunk_scan = [&](int v) -> int {
    s_scan[t] = v;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {
      const int x = t >= o ? s_scan[t - o] : 0;
      __syncthreads();
      s_scan[t] += x;
      __syncthreads();
    }
    const int excl = s_base + s_scan[t] - v;
    __syncthreads();
    if (t == (int)blockDim.x - 1) s_base += s_scan[t];
    __syncthreads();
    return excl;
  };
  long long kept = 0;
  for (int p = t; p < P; p += blockDim.x) kept += n_keep[p];
  const long long total_kept = block_sum(kept);
  const bool do_split = mixup > total_kept;
  if (do_split) {
    const long long budget = (long long)mixup - P;   // the greedy starts from one component per pdf
    auto max_splits = [&](double occ) -> int {        // splits allowed: from k to k+1 components while (k+1)*min_count < occ
      int nmax = (int)floor(occ / min_count);
      while (nmax > 0 && (double)nmax * min_count >= occ) nmax--;
      nmax = nmax < 1 ? 1 : (nmax > MAXC ? MAXC : nmax);
      return nmax - 1;
    };
    auto count = [&](double theta) -> long long {
      long long c = 0;
      for (int p = t; p < P; p += blockDim.x) { const double so = state_occ[p]; c += splits_at_least(pow(so > 0.0 ? so : 0.0, power), max_splits(so), theta); }
      return block_sum(c);
    };
    // bisection over the bit patterns of positive doubles (monotone in the value): the largest theta with count(theta) >= budget
    unsigned long long lo = 1ull, hi = 0x7FEFFFFFFFFFFFFFull;   // count(lo) = every allowed split, count(hi) = 0
    const long long all = count(__longlong_as_double((long long)lo));
    if (all <= budget) {
      for (int p = t; p < P; p += blockDim.x) { const double so = state_occ[p]; n_new[p] = max(n_keep[p], 1 + (so > 0.0 ? max_splits(so) : 0)); }
    } else {
      while (hi - lo > 1) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (count(__longlong_as_double((long long)mid)) >= budget) lo = mid; else hi = mid;
      }
      const double theta = __longlong_as_double((long long)lo), above = __longlong_as_double((long long)hi);
      // every key > theta is taken (fewer than the budget); the keys equal to theta (at least the marginal one) fill the rest in pdf order
      const int carry = (int)(budget - count(above));
      if (t == 0) s_base = 0;
      __syncthreads();
      for (int p0 = 0; p0 < P; p0 += blockDim.x) {
        const int p = p0 + t;
        int tg = 0, ties = 0;
        if (p < P) {
          const double so = state_occ[p], c = pow(so > 0.0 ? so : 0.0, power);
          const int ms = max_splits(so);
          const int n_above = splits_at_least(c, ms, above);
          ties = splits_at_least(c, ms, theta) - n_above;
          tg = 1 + n_above;
        }
        const int before = chunk_scan(ties);
        if (p < P) n_new[p] = max(n_keep[p], tg + min(ties, max(0, carry - before)));
      }
    }
  } else {
    for (int p = t; p < P; p += blockDim.x) n_new[p] = n_keep[p];
  }
  __syncthreads();
  if (t == 0) s_base = 0;
  __syncthreads();
  long long split = 0;
  for (int p0 = 0; p0 < P; p0 += blockDim.x) {
    const int p = p0 + t;
    const int v = p < P ? n_new[p] : 0;
    if (p < P) split += v - n_keep[p];
    const int excl = chunk_scan(v);
    if (p < P) new_off[p] = excl;
  }
  const long long n_split = block_sum(split);
  if (t == 0) {
    new_off[P] = s_base;
    scal->removed = (int32_t)(G_old - total_kept);
    scal->split = (int32_t)n_split;
    scal->gauss_after = s_base;
  }
}

// counter-based standard normal: splitmix64 of (seed, pdf, component, dimension) -> two uniforms -> Box-Muller
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double randn(unsigned long long seed, int pdf, int comp, int d) {
  const unsigned long long a = mix64(seed ^ mix64(((unsigned long long)(unsigned)pdf << 32) | ((unsigned long long)(unsigned)comp << 8) | (unsigned)d));
  const unsigned long long b = mix64(a);
  const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0), u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

struct WriteArgs {
  int P, D, remove;
  const int32_t *pdf_off, *new_off, *n_keep;
  const uint8_t *flags;
  const double *state_occ, *occ, *mean_acc, *var_acc;
  const float *gc_old, *miv_old, *iv_old;
  float *w_new, *gc_new, *miv_new, *iv_new;
  double min_var, perturb;
  unsigned long long seed;
  MleScal *scal;
};

// ---- 3: per pdf (one CTA): new parameters of the kept Gaussians staged in shared memory (f64), objective change, mix-up split,
// then the pdf's new rows (f32 means_invvars / inv_vars, weights, gconsts computed from the f32 values as DiagGmm::ComputeGconsts does)
__global__ void __launch_bounds__(128)
mle_write_kernel(WriteArgs a) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int D = a.D, p = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double *s_mu = (double *)sm_raw, *s_var = s_mu + (size_t)MAXC * D;
  __shared__ double s_w[MAXC], s_red[4];
  __shared__ int s_src[MAXC], s_cnt[4], s_arg;
  const int o0 = a.pdf_off[p], n = a.pdf_off[p + 1] - o0, nb = a.new_off[p], nk = a.n_keep[p], nn = a.new_off[p + 1] - nb;
  // kept list in index order
  const int keep = t < n ? (a.flags[o0 + t] & 1) : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_cnt[warp] = __popc(bal);
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; w++) base += s_cnt[w];
  const int r_mine = base + __popc(bal & ((1u << lane) - 1));
  const double so = a.state_occ[p];
  if (keep) {
    s_src[r_mine] = t;
    const int upd = (a.flags[o0 + t] >> 1) & 1;
    const double prob = so > 0.0 ? a.occ[o0 + t] / so : 1.0 / n;
    s_w[r_mine] = (upd || !a.remove) ? prob : 1.0;   // an un-updated survivor is the only component left: its weight renormalises to 1
  }
  __syncthreads();
  // stage means / variances of the kept components
  double floored = 0.0;
  for (int i = t; i < nk * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D, m = o0 + s_src[r];
    double mu, var;
    if ((a.flags[m] >> 1) & 1) {
      const double oc = a.occ[m];
      mu = a.mean_acc[(size_t)m * D + d] / oc;
      var = a.var_acc[(size_t)m * D + d] / oc - mu * mu;
      if (var < a.min_var) { var = a.min_var; floored += 1.0; }
    } else {
      const double iv = (double)a.iv_old[(size_t)m * D + d];
      mu = (double)a.miv_old[(size_t)m * D + d] / iv;
      var = 1.0 / iv;
    }
    s_mu[(size_t)r * D + d] = mu; s_var[(size_t)r * D + d] = var;
  }
  // weights renormalised over the kept components (RemoveComponents(..., renorm_weights = true))
  double ws = t < nk ? s_w[t] : 0.0;
  ws = warp_sum_d(ws);
  if (lane == 0) s_red[warp] = ws;
  __syncthreads();
  const double wsum = s_red[0] + s_red[1] + s_red[2] + s_red[3];
  __syncthreads();
  if (t < nk) s_w[t] /= wsum;
  __syncthreads();
  // objective change of this pdf (only defined when no component was removed; evaluated before the mix-up split, like Kaldi)
  double impr = 0.0, cnt = t < n ? a.occ[o0 + t] : 0.0;   // the count covers every component, removed ones included
  if (t < nk) {
    const int m = o0 + s_src[t];
    if (nk == n) {
      double lg = 0.0, qd = 0.0, lin = 0.0, quad = 0.0;
      for (int d = 0; d < D; d++) {
        const float ivf = (float)(1.0 / s_var[(size_t)t * D + d]), mivf = (float)(s_mu[(size_t)t * D + d] / s_var[(size_t)t * D + d]);
        lg += log((double)ivf); qd += (double)mivf * (double)mivf / (double)ivf;
        lin += a.mean_acc[(size_t)m * D + d] * ((double)mivf - (double)a.miv_old[(size_t)m * D + d]);
        quad += a.var_acc[(size_t)m * D + d] * ((double)ivf - (double)a.iv_old[(size_t)m * D + d]);
      }
      const float gcf = (float)((double)logf((float)s_w[t]) - 0.5 * kLog2Pi * D + 0.5 * lg - 0.5 * qd);
      impr = a.occ[m] * ((double)gcf - (double)a.gc_old[m]) + lin - 0.5 * quad;
    }
  }
  impr = warp_sum_d(impr); cnt = warp_sum_d(cnt); floored = warp_sum_d(floored);
  if (lane == 0) {
    if (impr != 0.0) atomicAdd(&a.scal->objf_impr, impr);
    if (cnt != 0.0) atomicAdd(&a.scal->count, cnt);
    if (floored != 0.0) atomicAdd(&a.scal->floored, floored);
  }
  // mix-up: DiagGmm::Split -- the heaviest component (first maximum) is halved until the pdf has nn components
  for (int cur = nk; cur < nn; cur++) {
    __syncthreads();
    double v = t < cur ? s_w[t] : -1.0;
    int arg = t;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
      const int a2 = __shfl_xor_sync(0xffffffffu, arg, o);
      if (v2 > v || (v2 == v && a2 < arg)) { v = v2; arg = a2; }
    }
    if (lane == 0) { s_red[warp] = v; s_cnt[warp] = arg; }
    __syncthreads();
    if (t == 0) {
      double bv = s_red[0]; int ba = s_cnt[0];
      for (int w = 1; w < 4; w++) if (s_red[w] > bv || (s_red[w] == bv && s_cnt[w] < ba)) { bv = s_red[w]; ba = s_cnt[w]; }
      s_arg = ba;
      s_w[ba] = bv * 0.5; s_w[cur] = bv * 0.5;
    }
    __syncthreads();
    const int k = s_arg;
    for (int d = t; d < D; d += blockDim.x) {
      const double var = s_var[(size_t)k * D + d];
      const double r = randn(a.seed, p, cur, d) * sqrt(var) * a.perturb;
      s_mu[(size_t)cur * D + d] = s_mu[(size_t)k * D + d] + r;
      s_mu[(size_t)k * D + d] -= r;
      s_var[(size_t)cur * D + d] = var;
    }
  }
  __syncthreads();
  // new rows
  for (int i = t; i < nn * D; i += blockDim.x) {
    const int r = i / D, d = i - r * D;
    const double var = s_var[(size_t)r * D + d];
    a.miv_new[(size_t)(nb + r) * D + d] = (float)(s_mu[(size_t)r * D + d] / var);
    a.iv_new[(size_t)(nb + r) * D + d] = (float)(1.0 / var);
  }
  if (t < nn) {
    double lg = 0.0, qd = 0.0;
    for (int d = 0; d < D; d++) {
      const double var = s_var[(size_t)t * D + d];
      const float ivf = (float)(1.0 / var), mivf = (float)(s_mu[(size_t)t * D + d] / var);
      lg += log((double)ivf); qd += (double)mivf * (double)mivf / (double)ivf;
    }
    const float wf = (float)s_w[t];
    double gc = (double)logf(wf) - 0.5 * kLog2Pi * D + 0.5 * lg - 0.5 * qd;
    if (!isfinite(gc)) gc = -1.0e20;   // Kaldi: a NaN / infinite gconst becomes very negative
    a.w_new[nb + t] = wf;
    a.gc_new[nb + t] = (float)gc;
  }
}

// ---- 4: transitions (one thread per transition-state), TransitionModel::MleUpdate (non-shared)
__global__ void mle_trans_kernel(int n_tstates, const int32_t *__restrict__ first_tid, const double *__restrict__ stats, float *__restrict__ log_probs,
                                 double floor_p, double mincount, MleScal *__restrict__ scal) {
  const int ts = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (ts > n_tstates) return;
  const int a = first_tid[ts], n = first_tid[ts + 1] - a;
  if (n <= 1) return;
  double tot = 0.0;
  for (int k = 0; k < n; k++) tot += stats[a + k];
  atomicAdd(&scal->trans_count, tot);
  if (tot < mincount) return;
  double impr = 0.0;
  // renormalise, then floor, three times (the floor is the last step); MFA's topologies have at most 4 transitions per state,
  // states with more than NMAX keep their probabilities
  constexpr int NMAX = 32;
  double pr[NMAX];
  if (n > NMAX) return;
  for (int k = 0; k < n; k++) pr[k] = stats[a + k];
  for (int pass = 0; pass < 3; pass++) {
    double sum = 0.0;
    for (int k = 0; k < n; k++) sum += pr[k];
    for (int k = 0; k < n; k++) pr[k] = fmax(pr[k] / sum, floor_p);
  }
  for (int k = 0; k < n; k++) {
    const double c = stats[a + k], old = exp((double)log_probs[a + k]);
    if (c > 0.0 && old > 0.0 && pr[k] > 0.0) impr += c * (log(pr[k]) - log(old));
    log_probs[a + k] = (float)log(pr[k]);
  }
  atomicAdd(&scal->trans_impr, impr);
}

// per-tid AddTransitionProbs cost from (updated) log-probabilities: -(scaled log prob), hmm/hmm-utils.cc
__global__ void tid_cost_kernel(int n_tstates, const int32_t *__restrict__ first_tid, const int32_t *__restrict__ self_loop_tid,
                                const float *__restrict__ log_probs, float tscale, float slscale, float *__restrict__ tid_cost) {
  const int ts = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (ts > n_tstates) return;
  const int a = first_tid[ts], b = first_tid[ts + 1], sl = self_loop_tid[ts];
  // non-self-loop log-probability of the state: log(1 - p_selfloop), floored like TransitionModel::ComputeDerivedOfProbs
  float nsl = 0.0f;
  if (sl != 0) {
    double q = 1.0 - exp((double)log_probs[sl]);
    if (q <= 0.0) q = 1.0e-10;
    nsl = (float)log(q);
  }
  for (int tid = a; tid < b; tid++) {
    const float lp = log_probs[tid];
    float v;   // rounded products and sums, no fused multiply-add: the same float arithmetic as the host packing path
    if (tscale == slscale) v = __fmul_rn(lp, tscale);
    else if (tid == sl) v = __fmul_rn(slscale, lp);
    else v = __fadd_rn(__fmul_rn(slscale, nsl), __fmul_rn(tscale, __fsub_rn(lp, nsl)));
    tid_cost[tid] = -v;
  }
  if (ts == 1) tid_cost[0] = 0.0f;
}

}  // namespace

extern "C" {

int mfa_model_set_transitions(mfa_model *m, const mfa_trans_desc *d) {
  if (!m || !d || !d->tstate_first_tid || !d->self_loop_tid || !d->log_probs || d->num_tstates < 0) return set_error(MFA_ERR_INVALID, "bad argument");
  if (d->tstate_first_tid[d->num_tstates + 1] != m->num_tids + 1) return set_error(MFA_ERR_INVALID, "transition tables do not cover num_tids");
  CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t s = m->eng->stream;
  for (void **p : {(void **)&m->d_first_tid, (void **)&m->d_self_loop_tid, (void **)&m->d_log_probs, (void **)&m->d_tid_cost})
    if (*p) { CUDA_TRY(cudaStreamSynchronize(s)); CUDA_TRY(cudaFree(*p)); *p = nullptr; }
  m->num_tstates = d->num_tstates;
  const size_t n1 = (size_t)d->num_tstates + 2, n2 = (size_t)d->num_tstates + 1, n3 = (size_t)m->num_tids + 1;
  CUDA_TRY(cudaMalloc((void **)&m->d_first_tid, n1 * 4)); CUDA_TRY(cudaMalloc((void **)&m->d_self_loop_tid, n2 * 4));
  CUDA_TRY(cudaMalloc((void **)&m->d_log_probs, n3 * 4)); CUDA_TRY(cudaMalloc((void **)&m->d_tid_cost, n3 * 4));
  CUDA_TRY(cudaMemcpyAsync(m->d_first_tid, d->tstate_first_tid, n1 * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(m->d_self_loop_tid, d->self_loop_tid, n2 * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(m->d_log_probs, d->log_probs, n3 * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return MFA_OK;
}

int mfa_model_reserve(mfa_engine *e, mfa_model *m, int64_t max_gauss) {
  if (!e || !m || max_gauss < m->num_gauss) return set_error(MFA_ERR_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  cudaStream_t s = e->stream;
  const size_t cap = (size_t)max_gauss + (size_t)max_gauss / 8, D = (size_t)m->dim, P = (size_t)m->num_pdfs, G = (size_t)m->num_gauss;
  CUDA_TRY(cudaStreamSynchronize(s));
  if (m->sp_cap < cap || !m->sp_pdf_off) {
    for (void **p : {(void **)&m->sp_weights, (void **)&m->sp_gconsts, (void **)&m->sp_miv, (void **)&m->sp_iv, (void **)&m->sp_pdf_off})
      if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; }
    m->sp_cap = cap;
    CUDA_TRY(cudaMalloc((void **)&m->sp_weights, cap * 4)); CUDA_TRY(cudaMalloc((void **)&m->sp_gconsts, cap * 4));
    CUDA_TRY(cudaMalloc((void **)&m->sp_miv, cap * D * 4)); CUDA_TRY(cudaMalloc((void **)&m->sp_iv, cap * D * 4));
    CUDA_TRY(cudaMalloc((void **)&m->sp_pdf_off, (P + 1) * 4));
  }
  if ((m->par_cap ? m->par_cap : G) < cap && m->d_weights) {
    // the current arrays move into new ones of the same capacity (after the first swap they become the spare set)
    float *w, *gc, *miv, *iv;
    CUDA_TRY(cudaMalloc((void **)&w, cap * 4)); CUDA_TRY(cudaMalloc((void **)&gc, cap * 4));
    CUDA_TRY(cudaMalloc((void **)&miv, cap * D * 4)); CUDA_TRY(cudaMalloc((void **)&iv, cap * D * 4));
    CUDA_TRY(cudaMemcpyAsync(w, m->d_weights, G * 4, cudaMemcpyDeviceToDevice, s)); CUDA_TRY(cudaMemcpyAsync(gc, m->d_gconsts, G * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(miv, m->d_miv, G * D * 4, cudaMemcpyDeviceToDevice, s)); CUDA_TRY(cudaMemcpyAsync(iv, m->d_iv, G * D * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (void *p : {(void *)m->d_weights, (void *)m->d_gconsts, (void *)m->d_miv, (void *)m->d_iv}) CUDA_TRY(cudaFree(p));
    m->d_weights = w; m->d_gconsts = gc; m->d_miv = miv; m->d_iv = iv; m->par_cap = cap;
  }
  // accumulator block for `cap` Gaussians
  const size_t acc_bytes = (cap * (1 + 2 * D) + (size_t)m->num_tids + 1 + 2) * sizeof(double);
  const bool had = m->d_acc != nullptr;
  if (!had || m->acc_cap_bytes < acc_bytes) {
    if (had) return MFA_OK;   // live accumulators: leave them alone, the next mfa_acc_zero after the M-step sizes the block
    MFA_TRY(m->acc_take(acc_bytes));
    m->acc_spare = m->d_acc; m->d_acc = nullptr;
  }
  return MFA_OK;
}

int mfa_model_mle_update(mfa_engine *e, mfa_model *m, const mfa_mle_opts *o, mfa_mle_result *res) {
  if (!e || !m || !o) return set_error(MFA_ERR_INVALID, "null argument");
  if (!m->d_acc) return set_error(MFA_ERR_INVALID, "no accumulators: run mfa_acc_zero / mfa_acc_stats first");
  if (o->update_transitions && !m->d_first_tid) return set_error(MFA_ERR_INVALID, "update_transitions needs mfa_model_set_transitions");
  CUDA_TRY(cudaSetDevice(e->device));
  CallScope scope(e);
  cudaStream_t s = e->stream;
  const int P = m->num_pdfs, G = m->num_gauss, D = m->dim, NT = m->num_tids;
  const int mixup = o->mixup > 0 ? o->mixup : 0;
  // new number of Gaussians = sum over pdfs of max(kept, mix-up target) <= kept + sum of targets <= G + mixup
  const size_t Gcap = (size_t)G + (size_t)mixup + 1;
  // scratch: flags[G] | state_occ[P] f64 | n_keep[P] | n_new[P] | new_off[P+1] | scalars
  uint8_t *scr;
  const size_t off_occ = ((size_t)G + 15) & ~(size_t)15, off_keep = off_occ + (size_t)P * 8, off_new = off_keep + (size_t)P * 4,
               off_off = off_new + (size_t)P * 4, off_scal = (off_off + ((size_t)P + 1) * 4 + 15) & ~(size_t)15, total = off_scal + sizeof(MleScal);
  MFA_TRY(e->getT<uint8_t>(DB_MLE, total, &scr));
  uint8_t *flags = scr; double *state_occ = (double *)(scr + off_occ); int32_t *n_keep = (int32_t *)(scr + off_keep), *n_new = (int32_t *)(scr + off_new),
          *new_off = (int32_t *)(scr + off_off); MleScal *scal = (MleScal *)(scr + off_scal);
  CUDA_TRY(cudaMemsetAsync(scal, 0, sizeof(MleScal), s));
  // the new parameters go into the spare set (allocated by capacity, kept across iterations), then the sets swap
  if (m->sp_cap < Gcap || !m->sp_pdf_off) {
    CUDA_TRY(cudaStreamSynchronize(s));
    for (void **p : {(void **)&m->sp_weights, (void **)&m->sp_gconsts, (void **)&m->sp_miv, (void **)&m->sp_iv, (void **)&m->sp_pdf_off})
      if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; }
    m->sp_cap = Gcap + Gcap / 8;
    CUDA_TRY(cudaMalloc((void **)&m->sp_weights, m->sp_cap * 4)); CUDA_TRY(cudaMalloc((void **)&m->sp_gconsts, m->sp_cap * 4));
    CUDA_TRY(cudaMalloc((void **)&m->sp_miv, m->sp_cap * D * 4)); CUDA_TRY(cudaMalloc((void **)&m->sp_iv, m->sp_cap * D * 4));
    CUDA_TRY(cudaMalloc((void **)&m->sp_pdf_off, ((size_t)P + 1) * 4));
  }
  float *w_new = m->sp_weights, *gc_new = m->sp_gconsts, *miv_new = m->sp_miv, *iv_new = m->sp_iv;
  const double *occ = m->d_acc, *mean_acc = occ + G, *var_acc = mean_acc + (size_t)G * D, *trans = var_acc + (size_t)G * D, *tot = trans + (NT + 1);
  mle_pdf_kernel<<<(unsigned)((P + 7) / 8), 256, 0, s>>>(P, m->d_pdf_off, occ, o->min_gaussian_occupancy, o->min_gaussian_weight,
                                                          o->remove_low_count_gaussians, flags, state_occ, n_keep);
  mle_plan_kernel<<<1, 1024, 0, s>>>(P, G, state_occ, n_keep, mixup, (double)o->power, (double)o->min_count, n_new, new_off, scal);
  WriteArgs a{};
  a.P = P; a.D = D; a.remove = o->remove_low_count_gaussians; a.pdf_off = m->d_pdf_off; a.new_off = new_off; a.n_keep = n_keep; a.flags = flags;
  a.state_occ = state_occ; a.occ = occ; a.mean_acc = mean_acc; a.var_acc = var_acc; a.gc_old = m->d_gconsts; a.miv_old = m->d_miv; a.iv_old = m->d_iv;
  a.w_new = w_new; a.gc_new = gc_new; a.miv_new = miv_new; a.iv_new = iv_new; a.min_var = o->min_variance; a.perturb = o->perturb_factor;
  a.seed = o->seed; a.scal = scal;
  const size_t smem = (size_t)2 * MAXC * D * sizeof(double);
  CUDA_TRY(cudaFuncSetAttribute(mle_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mle_write_kernel<<<(unsigned)P, 128, smem, s>>>(a);
  if (o->update_transitions)
    mle_trans_kernel<<<(unsigned)((m->num_tstates + 127) / 128), 128, 0, s>>>(m->num_tstates, m->d_first_tid, trans, m->d_log_probs,
                                                                             (double)o->transition_floor, (double)o->transition_mincount, scal);
  e->launches += o->update_transitions ? 4 : 3;
  CUDA_TRY(cudaGetLastError());
  // results: scalars + the new layout (the host plans K2 tiles from it)
  std::vector<int32_t> h_off((size_t)P + 1);
  MleScal h{};
  CUDA_TRY(cudaMemcpyAsync(&h, scal, sizeof(MleScal), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(h_off.data(), new_off, ((size_t)P + 1) * 4, cudaMemcpyDeviceToHost, s));
  double h_tot[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(h_tot, tot, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (h.gauss_after <= 0 || (size_t)h.gauss_after > Gcap - 1 || h_off[P] != h.gauss_after) {
    return set_error(MFA_ERR_INVALID, "internal: M-step produced an inconsistent layout (" + std::to_string(h.gauss_after) + " Gaussians)");
  }
  // swap the model over to the new arrays; the old ones become the spare set, the consumed accumulator block is kept for the next pass
  CUDA_TRY(cudaMemcpyAsync(m->sp_pdf_off, new_off, ((size_t)P + 1) * 4, cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  {
    const size_t old_cap = m->par_cap ? m->par_cap : (size_t)G;
    std::swap(m->d_pdf_off, m->sp_pdf_off); std::swap(m->d_gconsts, m->sp_gconsts); std::swap(m->d_miv, m->sp_miv); std::swap(m->d_iv, m->sp_iv);
    std::swap(m->d_weights, m->sp_weights);
    m->par_cap = m->sp_cap; m->sp_cap = m->sp_weights ? old_cap : 0;
    if (!m->sp_weights) {   // the model was created without weights: that spare array must exist next time
      for (void **p : {(void **)&m->sp_gconsts, (void **)&m->sp_miv, (void **)&m->sp_iv, (void **)&m->sp_pdf_off}) if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; }
    }
    m->acc_spare = m->d_acc; m->d_acc = nullptr;
    if (m->sp_cap < m->par_cap) {
      // the set that just became the spare is the model's original (exact-size) allocation: bring it to the same capacity NOW, so that
      // all allocation of a training run happens inside its first M-step
      for (void **p : {(void **)&m->sp_weights, (void **)&m->sp_gconsts, (void **)&m->sp_miv, (void **)&m->sp_iv}) if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; }
      m->sp_cap = m->par_cap;
      CUDA_TRY(cudaMalloc((void **)&m->sp_weights, m->sp_cap * 4)); CUDA_TRY(cudaMalloc((void **)&m->sp_gconsts, m->sp_cap * 4));
      CUDA_TRY(cudaMalloc((void **)&m->sp_miv, m->sp_cap * D * 4)); CUDA_TRY(cudaMalloc((void **)&m->sp_iv, m->sp_cap * D * 4));
      if (!m->sp_pdf_off) CUDA_TRY(cudaMalloc((void **)&m->sp_pdf_off, ((size_t)P + 1) * 4));
    }
  }
  const bool layout_changed = h_off != m->h_pdf_off;
  m->h_pdf_off = std::move(h_off);
  m->num_gauss = h.gauss_after;
  m->host_stale = true;        // h_gconsts / h_miv / h_iv / h_weights are refreshed from the device on demand
  m->ffma_ready = false;
  MFA_TRY(m->layout_tiles());
  MFA_TRY(build_tc_device(m, layout_changed));
  if (res) {
    res->gmm_objf_impr = h.objf_impr; res->gmm_count = h.count; res->trans_objf_impr = h.trans_impr; res->trans_count = h.trans_count;
    res->tot_like = h_tot[0]; res->tot_frames = h_tot[1]; res->num_gauss_before = G; res->num_gauss_after = h.gauss_after;
    res->num_removed = h.removed; res->num_split = h.split; res->variance_floored = (int64_t)h.floored; res->layout_changed = layout_changed ? 1 : 0;
  }
  return MFA_OK;
}

int mfa_model_read(mfa_engine *e, mfa_model *m, int32_t *pdf_off, float *weights, float *gconsts, float *means_invvars, float *inv_vars,
                   float *log_probs) {
  if (!e || !m) return set_error(MFA_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  MFA_TRY(e->join_k3());
  cudaStream_t s = e->stream;
  const size_t G = (size_t)m->num_gauss, D = (size_t)m->dim;
  if (pdf_off) memcpy(pdf_off, m->h_pdf_off.data(), ((size_t)m->num_pdfs + 1) * 4);
  if (weights) {
    if (!m->d_weights) return set_error(MFA_ERR_INVALID, "the model was created without weights");
    CUDA_TRY(cudaMemcpyAsync(weights, m->d_weights, G * 4, cudaMemcpyDeviceToHost, s));
  }
  if (gconsts) CUDA_TRY(cudaMemcpyAsync(gconsts, m->d_gconsts, G * 4, cudaMemcpyDeviceToHost, s));
  if (means_invvars) CUDA_TRY(cudaMemcpyAsync(means_invvars, m->d_miv, G * D * 4, cudaMemcpyDeviceToHost, s));
  if (inv_vars) CUDA_TRY(cudaMemcpyAsync(inv_vars, m->d_iv, G * D * 4, cudaMemcpyDeviceToHost, s));
  if (log_probs) {
    if (!m->d_log_probs) return set_error(MFA_ERR_INVALID, "no transition tables: call mfa_model_set_transitions first");
    CUDA_TRY(cudaMemcpyAsync(log_probs, m->d_log_probs, ((size_t)m->num_tids + 1) * 4, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return MFA_OK;
}

int mfa_model_num_gauss(const mfa_model *m) { return m ? m->num_gauss : 0; }

int mfa_graphs_set_transitions(mfa_engine *e, mfa_graphs *g, mfa_model *m, float transition_scale, float self_loop_scale) {
  if (!e || !g || !m) return set_error(MFA_ERR_INVALID, "null argument");
  if (!m->d_first_tid) return set_error(MFA_ERR_INVALID, "no transition tables: call mfa_model_set_transitions first");
  if (g->num_tids != m->num_tids) return set_error(MFA_ERR_INVALID, "graphs were packed for a different transition model");
  CUDA_TRY(cudaSetDevice(e->device));
  MFA_TRY(e->join_k3());   // a Viterbi launch in flight reads the arc weights this rewrites
  MFA_TRY(upload_graphs(e, g));
  tid_cost_kernel<<<(unsigned)((m->num_tstates + 127) / 128), 128, 0, e->stream>>>(m->num_tstates, m->d_first_tid, m->d_self_loop_tid, m->d_log_probs,
                                                                                  transition_scale, self_loop_scale, m->d_tid_cost);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return refold_graphs(e, g, m->d_tid_cost);
}

}  // extern "C"
