// mfcc.cu -- K1: batched MFCC (framing, DC removal, pre-emphasis, Povey window, real FFT, power spectrum,
// mel filterbank, log, DCT-II, lifter) + per-speaker CMVN statistics.
//
// Replaces kalpy MfccComputer.compute_mfccs_for_export (reference call sites: montreal_forced_aligner/
// corpus/features.py:235, online/alignment.py:83) and CmvnComputer (corpus/acoustic_corpus.py:1336).
// Algorithm per SURVEY.md A.2 (Kaldi feat/feature-window.cc, mel-computations.cc, feature-mfcc.cc).
//
// Mapping: one warp per frame; each warp walks a contiguous range of frames so the utterance lookup is
// amortised and neighbouring frames re-read their 60 % overlapping samples from L1/L2.
//   * mfcc512_kernel (the path MFA's defaults take: 25 ms at 16 kHz -> 512-point real FFT, <= 32 mel bins, <= 16 cepstra):
//     the 256-point complex FFT runs in REGISTERS, 8 points per lane: radix-8 over the in-lane index, twiddle, padded
//     shared-memory transpose, radix-8, twiddle, transpose, radix-4 (256 = 8 x 8 x 4); the real-FFT untangling works on
//     (k, 256-k) pairs and goes straight to the power spectrum; mel taps are split into <= 32 balanced chunks, one per lane;
//     the DCT is split over two half-warps.  ~600 warp instructions per frame instead of ~2 700.
//   * mfcc_kernel (any other geometry): 256-point complex Stockham radix-2 FFT in the warp's shared-memory ping-pong buffers.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "cuda_internal.cuh"

using namespace mfa;

namespace {

struct MfccTables {  // offsets (in floats) into one device blob
  int N, NP, NB, shift, nbins, nceps, log2nb;
  int off_window, off_tw, off_ptw, off_melw, off_melfirst, off_mellen, off_meloff, off_dct, off_lift, total;
  // fast path (mfcc512_kernel): per-lane mel chunks {first power bin, weight offset (floats, from the blob start), taps, mel bin},
  // per-bin {first chunk, number of chunks}; fast = 0 when the geometry does not fit
  int fast, off_chunk, off_binchunk, chunk_max;
  // fixed-trip-count views for the fast path: per lane kMelTaps chunk weights (zero padded), per (coefficient, half) kDctTaps DCT weights
  int off_cw, off_dct2, off_lift2;   // (the fast kernel stages only [off_melw, total) in shared memory)
};
constexpr int kMelTaps = 24;   // >= the largest chunk of MFA's 23-bin geometry (23 taps); geometries beyond it keep the runtime-length loop
constexpr int kDctTaps = 16;   // >= half the mel bins (fast path: <= 32 bins)
constexpr int kDctStride = 20; // row stride of the padded DCT table: 80 bytes keep the lanes' 16-byte reads on different banks

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
  // fast-path tables: split every bin's taps into chunks of at most C taps, C minimal such that <= 32 chunks result
  t.fast = 0; t.chunk_max = 0;
  t.off_chunk = (t.total + 3) / 4 * 4; t.off_binchunk = t.off_chunk + 4 * 32; t.off_cw = t.off_binchunk + 2 * 32;
  t.off_dct2 = t.off_cw + 32 * kMelTaps; t.off_lift2 = t.off_dct2 + 16 * 2 * kDctStride; t.total = t.off_lift2 + 16;
  std::vector<int> chunk(4 * 32, 0), binchunk(2 * 32, 0);
  if (t.NP == 512 && t.nbins <= 32 && t.nceps <= 16 && t.N >= 64 && t.N <= 512) {
    int C = 1;
    for (;; C++) { int n = 0; for (int b = 0; b < t.nbins; b++) n += (len[b] + C - 1) / C; if (n <= 32) break; }
    int lane = 0;
    for (int b = 0; b < t.nbins; b++) {
      int nc = (len[b] + C - 1) / C, per = (len[b] + nc - 1) / nc;   // even split inside the bin
      binchunk[2 * b] = lane; binchunk[2 * b + 1] = nc;
      for (int c = 0, done = 0; c < nc; c++, lane++) {
        int cnt = std::min(per, len[b] - done);
        chunk[4 * lane] = first[b] + done; chunk[4 * lane + 1] = t.off_melw + woff[b] + done; chunk[4 * lane + 2] = cnt; chunk[4 * lane + 3] = b;
        t.chunk_max = std::max(t.chunk_max, cnt);
        done += cnt;
      }
    }
    t.fast = 1;
  }
  blob.assign(t.total, 0.0f);
  memcpy(&blob[t.off_chunk], chunk.data(), chunk.size() * 4);
  memcpy(&blob[t.off_binchunk], binchunk.data(), binchunk.size() * 4);
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
  if (t.fast) {
    // the same weights again, laid out for loops with a compile-time trip count (padding taps are exact zeros: they add +0 to a finite sum)
    for (int lane = 0; lane < 32; lane++)
      for (int i = 0; i < chunk[4 * lane + 2] && i < kMelTaps; i++) blob[t.off_cw + lane * kMelTaps + i] = blob[chunk[4 * lane + 1] + i];
    const int half = (t.nbins + 1) / 2;
    for (int c = 0; c < t.nceps; c++)
      for (int h = 0; h < 2; h++)
        for (int j = 0; j < half && h * half + j < t.nbins; j++) blob[t.off_dct2 + (c * 2 + h) * kDctStride + j] = blob[t.off_dct + c * t.nbins + h * half + j];
    for (int c = 0; c < t.nceps; c++) blob[t.off_lift2 + c] = blob[t.off_lift + c];
  }
  return MFA_OK;
}

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32)
mfcc_kernel(MfccTables t, const float *__restrict__ tab, const int16_t *__restrict__ pcm, const int64_t *__restrict__ sample_off,
            const int64_t *__restrict__ frame_off, int n_utts, int64_t frame_base, int64_t n_frames, int64_t frames_per_warp, float *__restrict__ out,
            float preemph, int snip_edges, int remove_dc, int use_energy, int raw_energy, float log_energy_floor) {
  extern __shared__ float smem[];
  float *stab = smem;                                   // blob[0, off_chunk): everything but the fast path's tables
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *wbuf = smem + ((t.off_chunk + 3) / 4 * 4) + warp * (4 * t.NB + 4);  // two float2[NB] buffers per warp
  for (int i = threadIdx.x; i < t.off_chunk; i += blockDim.x) stab[i] = tab[i];
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

// ------------------------------------------------------------------------------------------------ fast path (512-point frames)
// in-place 8-point forward DFT (W = exp(-2 pi i / 8)), natural order in and out
__device__ __forceinline__ void dft4(float c0r, float c0i, float c1r, float c1i, float c2r, float c2i, float c3r, float c3i,
                                     float &y0r, float &y0i, float &y1r, float &y1i, float &y2r, float &y2i, float &y3r, float &y3i) {
  const float e0r = c0r + c2r, e0i = c0i + c2i, e1r = c0r - c2r, e1i = c0i - c2i;
  const float o0r = c1r + c3r, o0i = c1i + c3i, o1r = c1i - c3i, o1i = c3r - c1r;   // (c1 - c3) * (-i)
  y0r = e0r + o0r; y0i = e0i + o0i; y2r = e0r - o0r; y2i = e0i - o0i;
  y1r = e1r + o1r; y1i = e1i + o1i; y3r = e1r - o1r; y3i = e1i - o1i;
}
__device__ __forceinline__ void dft8(float (&xr)[8], float (&xi)[8]) {
  const float h = 0.70710678118654752f;
  float ar[4], ai[4], br[4], bi[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { ar[i] = xr[i] + xr[i + 4]; ai[i] = xi[i] + xi[i + 4]; br[i] = xr[i] - xr[i + 4]; bi[i] = xi[i] - xi[i + 4]; }
  float t;
  t = br[1]; br[1] = (t + bi[1]) * h; bi[1] = (bi[1] - t) * h;          // * (1 - i) / sqrt 2
  t = br[2]; br[2] = bi[2]; bi[2] = -t;                                  // * -i
  t = br[3]; br[3] = (bi[3] - t) * h; bi[3] = -(bi[3] + t) * h;         // * (-1 - i) / sqrt 2
  dft4(ar[0], ai[0], ar[1], ai[1], ar[2], ai[2], ar[3], ai[3], xr[0], xi[0], xr[2], xi[2], xr[4], xi[4], xr[6], xi[6]);
  dft4(br[0], bi[0], br[1], bi[1], br[2], bi[2], br[3], bi[3], xr[1], xi[1], xr[3], xi[3], xr[5], xi[5], xr[7], xi[7]);
}

constexpr int kFW = 4;                 // warps per CTA
constexpr int kT1 = 34;                // row stride (float2) of the first transpose [k1][m2]: 2 wavefronts per 64-bit access
constexpr int kT2 = 9;                 // row stride (float2) of the second transpose [lane][s]
constexpr int kTB = 32 * kT2;          // float2 entries of the transpose / spectrum buffer: max(8 * kT1, 32 * kT2, 256 + 16)
static_assert(kTB >= 8 * kT1 && kTB >= 256 + 16, "transpose buffer too small");
constexpr int kPW = 260 + kMelTaps;    // power spectrum [257] + zero padding read by the fixed-length mel loop
constexpr int kWarpFloats = 2 * kTB + kPW + 32 + 32;   // transpose / spectrum buffer, power spectrum, chunk sums, log-mel

// int16 -> float without the conversion unit (I2F runs on the quarter-rate XU pipe and was 24 % of this kernel's stall samples):
// 2^23 + 32768 + v is exactly representable, so the sample is planted in the mantissa of 2^23 and the bias subtracted -- exact.
__device__ __forceinline__ float s16_to_float(uint32_t u16) { return __uint_as_float(0x4B000000u | (u16 ^ 0x8000u)) - 8421376.0f; }

__global__ void __launch_bounds__(kFW * 32, 6)
mfcc512_kernel(MfccTables t, const float *__restrict__ tab, const int16_t *__restrict__ pcm, const int64_t *__restrict__ sample_off,
               const int64_t *__restrict__ frame_off, int n_utts, int64_t frame_base, int64_t n_frames, int64_t frames_per_warp, float *__restrict__ out,
               float preemph, int snip_edges, int remove_dc, int use_energy, int raw_energy, float log_energy_floor) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA-wide tables, laid out so that lane-indexed reads are conflict-free
  float2 *s_win = (float2 *)smem;                        // [8][32]  window[64 m1 + 2 lane + {0,1}]
  float2 *s_tw1 = s_win + 8 * 32;                        // [8][32]  W_256^(lane k1)
  float2 *s_tw2 = s_tw1 + 8 * 32;                        // [8][32]  W_32^((lane & 3) s)
  float *s_stage = (float *)(s_tw2 + 8 * 32);            // blob[off_melw, total): mel weights, chunk tables, padded mel / DCT weights, lifter
  const int n_stage = t.total - t.off_melw;
  const float *s_tab = s_stage - t.off_melw;             // so that blob offsets index it directly (only offsets >= off_melw are staged)
  float *wbase = s_stage + ((n_stage + 3) & ~3) + warp * kWarpFloats;
  float2 *tb = (float2 *)wbase;                          // kTB float2: both transposes and the spectrum Z (idx(k) = k + 4 (k >> 6))
  float *pw = wbase + 2 * kTB;                       // [257] power spectrum
  float *part = pw + kPW;                                // [32] per-chunk mel sums
  float *melv = part + 32;                               // [32] log mel energies
  const float2 *g_tw = (const float2 *)(tab + t.off_tw), *g_ptw = (const float2 *)(tab + t.off_ptw);
  for (int i = threadIdx.x; i < n_stage; i += blockDim.x) s_stage[i] = tab[t.off_melw + i];
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) {
    const int a = i >> 5, l = i & 31, n = 64 * a + 2 * l;
    s_win[i] = make_float2(n < t.N ? tab[t.off_window + n] : 0.0f, n + 1 < t.N ? tab[t.off_window + n + 1] : 0.0f);
    s_tw1[i] = g_tw[l * a];
    s_tw2[i] = g_tw[8 * (l & 3) * a];
  }
  __syncthreads();
  for (int i = 257 + lane; i < kPW; i += 32) pw[i] = 0.0f;   // never written again: the padding taps of the mel loop read them
  melv[lane] = 0.0f;                                        // slots >= nbins stay zero (padding taps of the DCT loop)
  __syncwarp();
  const int4 *chunks = (const int4 *)(s_tab + t.off_chunk);
  const int2 *binchunk = (const int2 *)(s_tab + t.off_binchunk);
  const float *lift = s_tab + t.off_lift2;
  // W_512^k for this lane's pairs k = lane + 32 i
  float2 ptw[4];
#pragma unroll
  for (int i = 0; i < 4; i++) ptw[i] = g_ptw[lane + 32 * i];
  const int4 my_chunk = chunks[lane];
  const int2 my_bin = lane < t.nbins ? binchunk[lane] : make_int2(0, 0);
  const int N = t.N, nbins = t.nbins, nceps = t.nceps;
  const int dc_c = lane & 15, dc_h = lane >> 4, dc_half = (nbins + 1) >> 1;   // DCT: coefficient dc_c, bins [dc_h * dc_half, ...)

  const int64_t gw = (int64_t)blockIdx.x * kFW + warp;
  int64_t f0 = frame_base + gw * frames_per_warp, f1 = f0 + frames_per_warp;
  if (f1 > frame_base + n_frames) f1 = frame_base + n_frames;
  if (f0 >= f1) return;
  int lo = 0, hi = n_utts - 1;
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (frame_off[mid] <= f0) lo = mid; else hi = mid - 1; }
  int u = lo;
  int64_t u_f0 = frame_off[u], u_f1 = frame_off[u + 1], u_s0 = sample_off[u], u_n = sample_off[u + 1] - u_s0;
  // the next frame's samples (same utterance, interior, 4-byte aligned, even window) are fetched one frame ahead: the load latency
  // at the top of a frame was a quarter of this kernel's stall samples
  uint32_t nx[8];
  bool have_nx = false;
  for (int64_t f = f0; f < f1; f++) {
    while (f >= u_f1) { u++; u_f0 = u_f1; u_f1 = frame_off[u + 1]; u_s0 = sample_off[u]; u_n = sample_off[u + 1] - u_s0; }
    const int64_t fi = f - u_f0;
    const int64_t start = snip_edges ? fi * t.shift : fi * t.shift + t.shift / 2 - N / 2;
    // ---- samples: lane holds x[64 a + 2 lane + {0,1}], a = 0..7 (128 contiguous bytes per load across the warp)
    float x[16];
    const int16_t *base = pcm + u_s0 + start;
    if (have_nx) {   // prefetched while the previous frame was being transformed
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int n = 64 * a + 2 * lane;
        x[2 * a] = n + 1 < N ? s16_to_float(nx[a] & 0xFFFFu) : 0.0f;
        x[2 * a + 1] = n + 1 < N ? s16_to_float(nx[a] >> 16) : 0.0f;
      }
    } else if (start >= 0 && start + N <= u_n) {
      if ((((size_t)base) & 3) == 0) {
#pragma unroll
        for (int a = 0; a < 8; a++) {
          const int n = 64 * a + 2 * lane;
          x[2 * a] = 0.0f; x[2 * a + 1] = 0.0f;
          if (n + 1 < N) { const uint32_t v = __ldg((const uint32_t *)(base + n)); x[2 * a] = s16_to_float(v & 0xFFFFu); x[2 * a + 1] = s16_to_float(v >> 16); }
          else if (n < N) x[2 * a] = s16_to_float((uint16_t)__ldg(base + n));
        }
      } else {
#pragma unroll
        for (int a = 0; a < 8; a++) {
          const int n = 64 * a + 2 * lane;
          x[2 * a] = n < N ? s16_to_float((uint16_t)__ldg(base + n)) : 0.0f;
          x[2 * a + 1] = n + 1 < N ? s16_to_float((uint16_t)__ldg(base + n + 1)) : 0.0f;
        }
      }
    } else {   // frame overlaps an utterance edge (snip_edges = false): reflected indices
#pragma unroll
      for (int a = 0; a < 8; a++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int n = 64 * a + 2 * lane + e;
          float v = 0.0f;
          if (n < N) {
            int64_t k = start + n;
            while (k < 0 || k >= u_n) k = (k < 0) ? -k - 1 : 2 * u_n - 1 - k;
            v = (float)pcm[u_s0 + k];
          }
          x[2 * a + e] = v;
        }
    }
    have_nx = false;
    if ((N & 1) == 0 && f + 1 < f1 && f + 1 < u_f1) {
      const int64_t start1 = start + t.shift;
      const int16_t *base1 = base + t.shift;
      if (start1 >= 0 && start1 + N <= u_n && (((size_t)base1) & 3) == 0) {
#pragma unroll
        for (int a = 0; a < 8; a++) { const int n = 64 * a + 2 * lane; nx[a] = n + 1 < N ? __ldg((const uint32_t *)(base1 + n)) : 0u; }
        have_nx = true;
      }
    }
    if (remove_dc) {
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; i++) sum += x[i];
#pragma unroll
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / (float)N;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int n = 64 * a + 2 * lane;
        if (n < N) x[2 * a] -= mean;
        if (n + 1 < N) x[2 * a + 1] -= mean;
      }
    }
    float energy = 0.0f;
    if (use_energy && raw_energy) {
#pragma unroll
      for (int i = 0; i < 16; i++) energy += x[i] * x[i];
    }
    // ---- pre-emphasis + window -> z[32 a + lane] = y[64 a + 2 lane] + i y[64 a + 2 lane + 1]
    float zr[8], zi[8];
    {
      float rot[8];
#pragma unroll
      for (int a = 0; a < 8; a++) rot[a] = __shfl_sync(0xffffffffu, x[2 * a + 1], (lane + 31) & 31);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const float xm1 = lane == 0 ? (a > 0 ? rot[a > 0 ? a - 1 : 0] : x[0]) : rot[a];
        const float2 w = s_win[a * 32 + lane];
        zr[a] = (x[2 * a] - preemph * xm1) * w.x;
        zi[a] = (x[2 * a + 1] - preemph * x[2 * a]) * w.y;
      }
    }
    if (use_energy && !raw_energy) {
#pragma unroll
      for (int a = 0; a < 8; a++) energy += zr[a] * zr[a] + zi[a] * zi[a];
    }
    // ---- 256-point complex FFT, k = k1 + 8 (s + 8 t)
    dft8(zr, zi);                                            // over a -> k1
#pragma unroll
    for (int k1 = 1; k1 < 8; k1++) {
      const float2 w = s_tw1[k1 * 32 + lane];
      const float r = zr[k1] * w.x - zi[k1] * w.y, i = zr[k1] * w.y + zi[k1] * w.x;
      zr[k1] = r; zi[k1] = i;
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; k1++) tb[k1 * kT1 + lane] = make_float2(zr[k1], zi[k1]);
    __syncwarp();
    {
      const int k1 = lane >> 2, j = lane & 3;                // lane (k1, j) takes m2 = 4 r + j
#pragma unroll
      for (int r = 0; r < 8; r++) { const float2 v = tb[k1 * kT1 + 4 * r + j]; zr[r] = v.x; zi[r] = v.y; }
    }
    __syncwarp();
    dft8(zr, zi);                                            // over r -> s
#pragma unroll
    for (int sx = 1; sx < 8; sx++) {
      const float2 w = s_tw2[sx * 32 + lane];
      const float r = zr[sx] * w.x - zi[sx] * w.y, i = zr[sx] * w.y + zi[sx] * w.x;
      zr[sx] = r; zi[sx] = i;
    }
#pragma unroll
    for (int sx = 0; sx < 8; sx++) tb[lane * kT2 + sx] = make_float2(zr[sx], zi[sx]);
    __syncwarp();
    {
      const int k1 = lane & 7, g = lane >> 3;                // lane (k1, g) finishes s = g and s = g + 4: 4-point DFT over j -> t
      float2 c[2][4];
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) c[e][j] = tb[(k1 * 4 + j) * kT2 + g + 4 * e];
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 2; e++) {
        float yr[4], yi[4];
        dft4(c[e][0].x, c[e][0].y, c[e][1].x, c[e][1].y, c[e][2].x, c[e][2].y, c[e][3].x, c[e][3].y, yr[0], yi[0], yr[1], yi[1], yr[2], yi[2], yr[3], yi[3]);
#pragma unroll
        for (int tt = 0; tt < 4; tt++) tb[k1 + 8 * (g + 4 * e) + 68 * tt] = make_float2(yr[tt], yi[tt]);   // idx(k), k = k1 + 8 s + 64 t
      }
    }
    __syncwarp();
    // ---- real-FFT untangling on pairs (k, 256 - k) -> power spectrum
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = lane + 32 * i;
      if (k == 0) {
        const float2 z = tb[0];
        pw[0] = (z.x + z.y) * (z.x + z.y); pw[256] = (z.x - z.y) * (z.x - z.y);
        const float2 zm = tb[128 + 4 * 2];
        pw[128] = zm.x * zm.x + zm.y * zm.y;
      } else {
        const int k2 = 256 - k;
        const float2 a = tb[k + 4 * (k >> 6)], b = tb[k2 + 4 * (k2 >> 6)];
        const float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);      // E = (Z[k] + conj Z[256-k]) / 2
        const float dr = 0.5f * (a.x - b.x), di = 0.5f * (a.y + b.y);      // D = (Z[k] - conj Z[256-k]) / 2;  O = D / i = (di, -dr)
        const float gr = di * ptw[i].x + dr * ptw[i].y, gi = di * ptw[i].y - dr * ptw[i].x;   // G = W_512^k O
        const float p1r = er + gr, p1i = ei + gi, p2r = er - gr, p2i = ei - gi;
        pw[k] = p1r * p1r + p1i * p1i;
        pw[k2] = p2r * p2r + p2i * p2i;
      }
    }
    __syncwarp();
    // ---- mel: one balanced chunk of taps per lane, chunks of a bin summed in order, log
    {
      // fixed trip count, weights zero-padded (the runtime-length loop cost 8 instructions per tap: 121 of the frame's 1 170)
      float e = 0.0f;
      const float *pp = pw + my_chunk.x;
      if (t.chunk_max <= kMelTaps) {
        const float4 *w4 = (const float4 *)(s_tab + t.off_cw + lane * kMelTaps);
#pragma unroll
        for (int q = 0; q < kMelTaps / 4; q++) {
          const float4 w = w4[q];
          e += w.x * pp[4 * q]; e += w.y * pp[4 * q + 1]; e += w.z * pp[4 * q + 2]; e += w.w * pp[4 * q + 3];
        }
      } else {
        const float *w = s_tab + my_chunk.y;
        for (int i = 0; i < my_chunk.z; i++) e += w[i] * pp[i];
      }
      part[lane] = e;
    }
    if (use_energy) {
#pragma unroll
      for (int o = 16; o; o >>= 1) energy += __shfl_xor_sync(0xffffffffu, energy, o);
    }
    __syncwarp();
    if (lane < nbins) {
      float e = part[my_bin.x];
      for (int c = 1; c < my_bin.y; c++) e += part[my_bin.x + c];
      melv[lane] = logf(fmaxf(e, FLT_EPSILON));
    }
    __syncwarp();
    // ---- DCT-II + lifter: lanes (c, half) take half of the bins each
    {
      float acc = 0.0f;
      if (dc_c < nceps) {
        const float4 *d4 = (const float4 *)(s_tab + t.off_dct2 + (dc_c * 2 + dc_h) * kDctStride);
        const float *mv = melv + dc_h * dc_half;
#pragma unroll
        for (int q = 0; q < kDctTaps / 4; q++) {
          const float4 d = d4[q];
          acc += d.x * mv[4 * q]; acc += d.y * mv[4 * q + 1]; acc += d.z * mv[4 * q + 2]; acc += d.w * mv[4 * q + 3];
        }
      }
      acc += __shfl_down_sync(0xffffffffu, acc, 16);
      if (lane < nceps) {
        acc *= lift[lane];
        if (use_energy && lane == 0) { const float le = logf(fmaxf(energy, FLT_EPSILON)); acc = fmaxf(le, log_energy_floor); }
        out[f * nceps + lane] = acc;
      }
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
  if (spk_utt_off[s] == spk_utt_off[s + 1]) return;   // speaker has no utterance in this call: its (zeroed or earlier) statistics stay
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
  const bool fast = t.fast && !e->cfg.mfcc_generic;
  if (fast) {
    const size_t smem = (size_t)(3 * 8 * 32 * 2 + ((t.total - t.off_melw + 3) & ~3) + kFW * kWarpFloats) * sizeof(float);
    CUDA_TRY(cudaFuncSetAttribute(mfcc512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t max_warps = (int64_t)e->sm_count * 6 * kFW * 4;   // ~4 waves of 6 resident CTAs per SM: the tables are rebuilt per CTA
    int64_t fpw = (n_frames + max_warps - 1) / max_warps;
    if (fpw < 8) fpw = 8;
    int64_t warps = (n_frames + fpw - 1) / fpw;
    int blocks = (int)((warps + kFW - 1) / kFW);
    float lef = (o->energy_floor > 0.0f) ? logf(o->energy_floor) : -INFINITY;
    mfcc512_kernel<<<blocks, kFW * 32, smem, e->stream>>>(t, d_tab, d_pcm, d_sample_off, d_frame_off, n_utts, frame_base, n_frames, fpw, d_out,
                                                             o->preemph_coeff, o->snip_edges, o->remove_dc_offset, o->use_energy, o->raw_energy, lef);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return MFA_OK;
  }
  size_t smem = ((t.off_chunk + 3) / 4 * 4 + kWarps * (4 * t.NB + 4)) * sizeof(float);
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
