// mfcc.cu -- K1: batched MFCC (framing, DC removal, pre-emphasis, Povey window, real FFT, power spectrum,
// mel filterbank, log, DCT-II, lifter) + per-speaker CMVN statistics.
//
// Replaces kalpy MfccComputer.compute_mfccs_for_export (reference call sites: montreal_forced_aligner/
// corpus/features.py:235, online/alignment.py:83) and CmvnComputer (corpus/acoustic_corpus.py:1336).
// Algorithm per SURVEY.md A.2 (Kaldi feat/feature-window.cc, mel-computations.cc, feature-mfcc.cc).
//
// Mapping: one warp per frame; each warp walks a contiguous range of frames so the utterance lookup is
// amortised and neighbouring frames re-read their 60 % overlapping samples from L1/L2.  The 512-point real
// FFT is a 256-point complex Stockham radix-2 FFT in the warp's private shared-memory ping-pong buffers.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {

struct MfccTables {  // offsets (in floats) into one device blob
  int N, NP, NB, shift, nbins, nceps, log2nb;
  int off_window, off_tw, off_ptw, off_melw, off_melfirst, off_mellen, off_meloff, off_dct, off_lift, total;
};

static int round_up_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
static float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

// Builds the constant tables on the host (f64 where Kaldi uses f64, f32 where it uses BaseFloat).
static int build_tables(const mfa_mfcc_opts *o, MfccTables &t, std::vector<float> &blob) {
  t.N = (int)(o->sample_frequency * 0.001f * o->frame_length_ms);
  t.shift = (int)(o->sample_frequency * 0.001f * o->frame_shift_ms);
  t.NP = round_up_pow2(t.N); t.NB = t.NP / 2; t.nbins = o->num_mel_bins; t.nceps = o->num_ceps;
  if (t.N < 2 || t.shift < 1) return set_error(MFA_ERR_INVALID, "bad frame length / shift");
  if (t.NB < 32 || t.NB > 1024) return set_error(MFA_ERR_UNSUPPORTED, "padded window size must be in 64..2048 samples");
  if (t.nbins < 3 || t.nbins > 128 || t.nceps < 1 || t.nceps > t.nbins || t.nceps > 32) return set_error(MFA_ERR_INVALID, "bad num_mel_bins / num_ceps");
  t.log2nb = 0; while ((1 << t.log2nb) < t.NB) t.log2nb++;
  int off = 0;
  auto take = [&](int n) { int r = off; off += (n + 3) / 4 * 4; return r; };
  t.off_window = take(t.N); t.off_tw = take(2 * t.NB); t.off_ptw = take(2 * t.NB);
  t.off_melfirst = take(t.nbins); t.off_mellen = take(t.nbins); t.off_meloff = take(t.nbins);
  t.off_dct = take(t.nceps * t.nbins); t.off_lift = take(t.nceps);
  t.off_melw = off;
  // mel weights first (variable length)
  std::vector<float> melw; std::vector<int> first(t.nbins), len(t.nbins), woff(t.nbins);
  float nyquist = 0.5f * o->sample_frequency;
  float high = o->high_freq > 0.0f ? o->high_freq : nyquist + o->high_freq;
  if (o->low_freq < 0.0f || o->low_freq >= nyquist || high <= 0.0f || high > nyquist || high <= o->low_freq) return set_error(MFA_ERR_INVALID, "bad low/high frequency");
  float bin_width = o->sample_frequency / t.NP;
  float mel_low = mel_scale(o->low_freq), mel_high = mel_scale(high);
  float mel_delta = (mel_high - mel_low) / (t.nbins + 1);
  for (int b = 0; b < t.nbins; b++) {
    float left = mel_low + b * mel_delta, center = mel_low + (b + 1) * mel_delta, right = mel_low + (b + 2) * mel_delta;
    first[b] = -1; int last = -1; std::vector<float> w(t.NB, 0.0f);
    for (int i = 0; i < t.NB; i++) {
      float mel = mel_scale(bin_width * i);
      if (mel > left && mel < right) {
        w[i] = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
        if (first[b] < 0) first[b] = i;
        last = i;
      }
    }
    if (first[b] < 0) return set_error(MFA_ERR_INVALID, "mel bin without FFT bins (too many mel bins)");
    len[b] = last - first[b] + 1; woff[b] = (int)melw.size();
    for (int i = first[b]; i <= last; i++) melw.push_back(w[i]);
  }
  t.total = t.off_melw + (int)melw.size();
  blob.assign(t.total, 0.0f);
  for (int i = 0; i < t.N; i++) blob[t.off_window + i] = (float)pow(0.5 - 0.5 * cos(2.0 * M_PI / (t.N - 1) * (double)i), 0.85);
  for (int k = 0; k < t.NB; k++) {
    blob[t.off_tw + 2 * k] = (float)cos(2.0 * M_PI * k / t.NB); blob[t.off_tw + 2 * k + 1] = (float)(-sin(2.0 * M_PI * k / t.NB));
    blob[t.off_ptw + 2 * k] = (float)cos(2.0 * M_PI * k / t.NP); blob[t.off_ptw + 2 * k + 1] = (float)(-sin(2.0 * M_PI * k / t.NP));
  }
  for (int b = 0; b < t.nbins; b++) {
    int v; v = first[b]; memcpy(&blob[t.off_melfirst + b], &v, 4); v = len[b]; memcpy(&blob[t.off_mellen + b], &v, 4);
    v = woff[b]; memcpy(&blob[t.off_meloff + b], &v, 4);
  }
  for (int k = 0; k < t.nceps; k++)
    for (int j = 0; j < t.nbins; j++)
      blob[t.off_dct + k * t.nbins + j] = (k == 0) ? (float)sqrt(1.0 / t.nbins) : (float)(sqrt(2.0 / t.nbins) * cos(M_PI / t.nbins * (j + 0.5) * k));
  for (int k = 0; k < t.nceps; k++)
    blob[t.off_lift + k] = (o->cepstral_lifter != 0.0f) ? (float)(1.0 + 0.5 * o->cepstral_lifter * sin(M_PI * k / o->cepstral_lifter)) : 1.0f;
  memcpy(&blob[t.off_melw], melw.data(), melw.size() * sizeof(float));
  return MFA_OK;
}

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32)
mfcc_kernel(MfccTables t, const float *__restrict__ tab, const int16_t *__restrict__ pcm, const int64_t *__restrict__ sample_off,
            const int64_t *__restrict__ frame_off, int n_utts, int64_t frame_base, int64_t n_frames, int64_t frames_per_warp, float *__restrict__ out,
            float preemph, int snip_edges, int remove_dc, int use_energy, int raw_energy, float log_energy_floor) {
  extern __shared__ float smem[];
  float *stab = smem;                                   // t.total floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *wbuf = smem + ((t.total + 3) / 4 * 4) + warp * (4 * t.NB + 4);  // two float2[NB] buffers per warp
  for (int i = threadIdx.x; i < t.total; i += blockDim.x) stab[i] = tab[i];
  __syncthreads();
  const float *window = stab + t.off_window;
  const float2 *tw = (const float2 *)(stab + t.off_tw), *ptw = (const float2 *)(stab + t.off_ptw);
  const int *melfirst = (const int *)(stab + t.off_melfirst), *mellen = (const int *)(stab + t.off_mellen), *meloff = (const int *)(stab + t.off_meloff);
  const float *melw = stab + t.off_melw, *dct = stab + t.off_dct, *lift = stab + t.off_lift;
  float2 *bufA = (float2 *)wbuf, *bufB = (float2 *)(wbuf + 2 * t.NB);

  const int64_t gw = (int64_t)blockIdx.x * kWarps + warp;
  int64_t f0 = frame_base + gw * frames_per_warp, f1 = f0 + frames_per_warp;
  if (f1 > frame_base + n_frames) f1 = frame_base + n_frames;
  if (f0 >= f1) return;
  // utterance of frame f0: largest u with frame_off[u] <= f0
  int lo = 0, hi = n_utts - 1;
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (frame_off[mid] <= f0) lo = mid; else hi = mid - 1; }
  int u = lo;
  int64_t u_f0 = frame_off[u], u_f1 = frame_off[u + 1], u_s0 = sample_off[u], u_n = sample_off[u + 1] - u_s0;
  const int N = t.N, NP = t.NP, NB = t.NB;
  for (int64_t f = f0; f < f1; f++) {
    while (f >= u_f1) { u++; u_f0 = u_f1; u_f1 = frame_off[u + 1]; u_s0 = sample_off[u]; u_n = sample_off[u + 1] - u_s0; }
    const int64_t fi = f - u_f0;
    const int64_t start = snip_edges ? fi * t.shift : fi * t.shift + t.shift / 2 - N / 2;
    float *xs = (float *)bufB;  // raw (DC-removed) samples
    float sum = 0.0f;
    for (int i = lane; i < N; i += 32) {
      int64_t k = start + i;
      while (k < 0 || k >= u_n) k = (k < 0) ? -k - 1 : 2 * u_n - 1 - k;
      float v = (float)pcm[u_s0 + k];
      xs[i] = v; sum += v;
    }
    if (remove_dc) {
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      float mean = sum / (float)N;
      for (int i = lane; i < N; i += 32) xs[i] -= mean;
    }
    __syncwarp();
    float energy = 0.0f;
    float *y = (float *)bufA;
    for (int i = lane; i < NP; i += 32) {
      float v = 0.0f;
      if (i < N) {
        float x = xs[i], xm1 = xs[i > 0 ? i - 1 : 0];
        if (use_energy && raw_energy) energy += x * x;
        v = (x - preemph * xm1) * window[i];
        if (use_energy && !raw_energy) energy += v * v;
      }
      y[i] = v;
    }
    __syncwarp();
    // Stockham radix-2: NB-point complex FFT of z[i] = y[2i] + i y[2i+1]
    float2 *in = bufA, *outb = bufB;
    for (int s = 0, Ns = 1; s < t.log2nb; s++, Ns <<= 1) {
      for (int j = lane; j < NB / 2; j += 32) {
        int k = j & (Ns - 1);
        float2 w = tw[k * (NB / (2 * Ns))];
        float2 a = in[j], b = in[j + NB / 2];
        float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
        int j0 = ((j - k) << 1) + k;
        outb[j0] = make_float2(a.x + bw.x, a.y + bw.y);
        outb[j0 + Ns] = make_float2(a.x - bw.x, a.y - bw.y);
      }
      __syncwarp();
      float2 *tmp = in; in = outb; outb = tmp;
    }
    // real-FFT post-processing -> power spectrum pw[0..NB] (written over the free buffer)
    float *pw = (float *)outb;
    for (int k = lane; k <= NB; k += 32) {
      float p;
      if (k == 0) { float2 z = in[0]; p = (z.x + z.y) * (z.x + z.y); }
      else if (k == NB) { float2 z = in[0]; p = (z.x - z.y) * (z.x - z.y); }
      else {
        float2 a = in[k], bc = in[NB - k];
        float br = bc.x, bi = -bc.y;
        float er = 0.5f * (a.x + br), ei = 0.5f * (a.y + bi);
        float dr = 0.5f * (a.x - br), di = 0.5f * (a.y - bi);
        float2 w = ptw[k];
        float orr = di, oi = -dr;
        float xr = er + (orr * w.x - oi * w.y), xi = ei + (orr * w.y + oi * w.x);
        p = xr * xr + xi * xi;
      }
      pw[k] = p;
    }
    __syncwarp();
    // mel energies (lane b handles bin b), log
    float *melv = (float *)in;  // FFT result no longer needed
    __syncwarp();
    for (int b = lane; b < t.nbins; b += 32) {
      const float *w = melw + meloff[b]; const float *p = pw + melfirst[b];
      float e = 0.0f;
      for (int i = 0, n = mellen[b]; i < n; i++) e += w[i] * p[i];
      melv[b] = logf(fmaxf(e, FLT_EPSILON));
    }
    if (use_energy) {
#pragma unroll
      for (int o = 16; o; o >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, o);
    }
    __syncwarp();
    if (lane < t.nceps) {
      float acc = 0.0f;
      const float *d = dct + lane * t.nbins;
      for (int j = 0; j < t.nbins; j++) acc += d[j] * melv[j];
      acc *= lift[lane];
      if (use_energy && lane == 0) { float le = logf(fmaxf(energy, FLT_EPSILON)); acc = fmaxf(le, log_energy_floor); }
      out[f * t.nceps + lane] = acc;
    }
    __syncwarp();
  }
}

// ---- CMVN statistics: per-utterance partial sums (f64), then per-speaker sums in utterance order (deterministic)
__global__ void cmvn_partial_kernel(const float *__restrict__ feats, int dim, const int64_t *__restrict__ frame_off, double *__restrict__ part) {
  const int u = blockIdx.x;
  const int64_t f0 = frame_off[u], T = frame_off[u + 1] - f0;
  const int lanes = blockDim.x / dim;  // frame lanes
  const int d = threadIdx.x % dim, fl = threadIdx.x / dim;
  double s = 0.0, s2 = 0.0;
  if (fl < lanes)
    for (int64_t t = fl; t < T; t += lanes) { double v = feats[(f0 + t) * dim + d]; s += v; s2 += v * v; }
  extern __shared__ double sh[];
  if (fl < lanes) { sh[(fl * dim + d) * 2] = s; sh[(fl * dim + d) * 2 + 1] = s2; }
  __syncthreads();
  if (threadIdx.x < dim) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; l++) { a += sh[(l * dim + threadIdx.x) * 2]; b += sh[(l * dim + threadIdx.x) * 2 + 1]; }
    part[((size_t)u * 2) * dim + threadIdx.x] = a;
    part[((size_t)u * 2 + 1) * dim + threadIdx.x] = b;
  }
}

__global__ void cmvn_reduce_kernel(const double *__restrict__ part, int dim, const int64_t *__restrict__ frame_off, const int32_t *__restrict__ spk_utt_off,
                                   const int32_t *__restrict__ spk_utts, int n_spk, double *__restrict__ stats) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_spk * (dim + 1)) return;
  int s = idx / (dim + 1), d = idx % (dim + 1);
  double a = 0.0, b = 0.0;
  for (int k = spk_utt_off[s]; k < spk_utt_off[s + 1]; k++) {
    int u = spk_utts[k];
    if (d < dim) { a += part[((size_t)u * 2) * dim + d]; b += part[((size_t)u * 2 + 1) * dim + d]; }
    else a += (double)(frame_off[u + 1] - frame_off[u]);
  }
  double *st = stats + (size_t)s * 2 * (dim + 1);
  st[d] = a; st[(dim + 1) + d] = (d < dim) ? b : 0.0;
}

}  // namespace

namespace mfa {

// d_sample_off / d_frame_off are the (global) offset arrays of the utterances [0, n_utts) handed in; the launch covers the frames
// [frame_base, frame_base + n_frames) of that numbering (pieces of a batch can be launched as their PCM arrives).
int launch_mfcc(mfa_engine *e, const mfa_mfcc_opts *o, const int16_t *d_pcm, const int64_t *d_sample_off, int32_t n_utts,
                const int64_t *d_frame_off, int64_t n_frames, float *d_out, int64_t frame_base) {
  if (n_frames == 0 || n_utts == 0) return MFA_OK;
  static_assert(sizeof(MfccTables) <= sizeof(e->mfcc_tab_desc), "table descriptor cache too small");
  MfccTables t;
  float *d_tab;
  if (e->mfcc_tab_valid && memcmp(&e->mfcc_tab_opts, o, sizeof(*o)) == 0) {
    memcpy(&t, e->mfcc_tab_desc, sizeof(t));
    d_tab = (float *)e->dev[DB_MFCC_TAB].p;
  } else {
    std::vector<float> blob;
    MFA_TRY(build_tables(o, t, blob));
    MFA_TRY(e->getT<float>(DB_MFCC_TAB, blob.size(), &d_tab));
    CUDA_TRY(cudaMemcpyAsync(d_tab, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));  // blob is a local
    memcpy(e->mfcc_tab_desc, &t, sizeof(t)); e->mfcc_tab_opts = *o; e->mfcc_tab_valid = true;
  }
  size_t smem = ((t.total + 3) / 4 * 4 + kWarps * (4 * t.NB + 4)) * sizeof(float);
  if (smem > e->smem_optin) return set_error(MFA_ERR_UNSUPPORTED, "MFCC tables exceed shared memory");
  CUDA_TRY(cudaFuncSetAttribute(mfcc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t max_warps = (int64_t)e->sm_count * 16 * kWarps;  // ~16 resident CTAs of 4 warps per SM
  int64_t fpw = (n_frames + max_warps - 1) / max_warps;
  if (fpw < 4) fpw = 4;
  int64_t warps = (n_frames + fpw - 1) / fpw;
  int blocks = (int)((warps + kWarps - 1) / kWarps);
  float lef = (o->energy_floor > 0.0f) ? logf(o->energy_floor) : -INFINITY;
  mfcc_kernel<<<blocks, kWarps * 32, smem, e->stream>>>(t, d_tab, d_pcm, d_sample_off, d_frame_off, n_utts, frame_base, n_frames, fpw, d_out,
                                                         o->preemph_coeff, o->snip_edges, o->remove_dc_offset, o->use_energy, o->raw_energy, lef);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

int launch_cmvn_stats(mfa_engine *e, const float *d_feats, int dim, const int64_t *d_frame_off, const int32_t *h_utt2spk, int32_t n_utts,
                      int32_t n_spk, double *d_stats) {
  if (n_utts == 0 || n_spk == 0) return MFA_OK;
  if (dim > 128) return set_error(MFA_ERR_UNSUPPORTED, "CMVN dim > 128");
  std::vector<int32_t> off(n_spk + 1, 0), utts(n_utts);
  for (int u = 0; u < n_utts; u++) { int s = h_utt2spk[u]; if (s < 0 || s >= n_spk) return set_error(MFA_ERR_INVALID, "utt2spk out of range"); off[s + 1]++; }
  for (int s = 0; s < n_spk; s++) off[s + 1] += off[s];
  { std::vector<int32_t> cur(off.begin(), off.end() - 1); for (int u = 0; u < n_utts; u++) utts[cur[h_utt2spk[u]]++] = u; }
  int32_t *d_off, *d_utts; double *d_part;
  MFA_TRY(e->upload(DB_SPK_UTT_OFF, off.data(), off.size(), &d_off));
  MFA_TRY(e->upload(DB_SPK_UTTS, utts.data(), utts.size(), &d_utts));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  MFA_TRY(e->getT<double>(DB_CMVN_PART, (size_t)n_utts * 2 * dim, &d_part));
  int threads = 256, lanes = threads / dim;
  cmvn_partial_kernel<<<n_utts, threads, (size_t)lanes * dim * 2 * sizeof(double), e->stream>>>(d_feats, dim, d_frame_off, d_part);
  e->launches++;
  int total = n_spk * (dim + 1);
  cmvn_reduce_kernel<<<(total + 127) / 128, 128, 0, e->stream>>>(d_part, dim, d_frame_off, d_off, d_utts, n_spk, d_stats);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return MFA_OK;
}

}  // namespace mfa

extern "C" int64_t mfa_mfcc_num_frames(const mfa_mfcc_opts *o, int64_t n) {
  int64_t shift = (int64_t)(o->sample_frequency * 0.001f * o->frame_shift_ms), len = (int64_t)(o->sample_frequency * 0.001f * o->frame_length_ms);
  if (o->snip_edges) return n < len ? 0 : 1 + (n - len) / shift;
  return (n + shift / 2) / shift;
}
