/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Scalar CPU restatement of the algorithms MFA's alignment hot path executes inside the
 * un-vendored third-party dependency kalpy 0.6.7 / Kaldi (not present under /root/reference):
 *   MFCC+CMVN -> deltas | splice+LDA -> fMLLR -> diagonal-GMM log-likelihoods -> FasterDecoder
 *   Viterbi with AlignUtteranceWrapper's retry -> GMM accumulator statistics; plus EqualAlign (monophone
 *   iteration 0) and fMLLR estimation (statistics + row-by-row update) between the two alignment passes.
 *
 * PARITY UNPINNED: the reference's tests hold no numeric golden vectors for this path and kalpy
 * cannot be installed offline, so this restatement follows the published Kaldi algorithms
 *  and is pinned only by (a) torchaudio.compliance.kaldi (MFCC, an
 * independent port), (b) closed-form checks (stored <GCONSTS> of the fixture models), and
 * (c) exhaustive-search cross-checks in tests/.
 *
 * Every function cites the reference call site it stands in for (paths relative to
 * /root/reference/montreal_forced_aligner) and the Kaldi source it restates.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * MFCC.  Call sites: corpus/features.py:235 (MfccFunction -> compute_mfccs_for_export),
 * online/alignment.py:83.  Restates Kaldi feat/feature-window.cc (ExtractWindow, ProcessWindow),
 * feat/mel-computations.cc (MelBanks, ComputeLifterCoeffs), matrix/matrix-functions (DCT),
 * feat/feature-mfcc.cc (MfccComputer::Compute).  SURVEY.md A.2.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float sample_frequency; /* 16000 */
  float frame_length_ms;  /* 25 */
  float frame_shift_ms;   /* 10 */
  float preemph_coeff;    /* 0.97 */
  float low_freq;         /* 20 */
  float high_freq;        /* 7800 (<=0: offset from nyquist) */
  float cepstral_lifter;  /* 22 */
  float energy_floor;     /* 0 */
  int32_t num_mel_bins;   /* 23 */
  int32_t num_ceps;       /* 13 */
  int32_t use_energy;     /* 0 */
  int32_t raw_energy;     /* 1 */
  int32_t snip_edges;     /* 1 */
  int32_t remove_dc_offset; /* 1 */
} orc_mfcc_opts;

static int round_up_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

static int win_shift(const orc_mfcc_opts *o) { return (int)(o->sample_frequency * 0.001f * o->frame_shift_ms); }
static int win_size(const orc_mfcc_opts *o) { return (int)(o->sample_frequency * 0.001f * o->frame_length_ms); }

ORC_API int64_t orc_mfcc_num_frames(const orc_mfcc_opts *o, int64_t n) {
  int64_t shift = win_shift(o), len = win_size(o);
  if (o->snip_edges) { if (n < len) return 0; return 1 + (n - len) / shift; }
  return (n + shift / 2) / shift;
}

/* in-place radix-2 complex FFT, n power of two (Kaldi uses split-radix; same transform) */
static void cfft(float *re, float *im, int n, const double *cs, const double *sn) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { float t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1, step = n / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < half; k++) {
        float wr = (float)cs[k * step], wi = (float)(-sn[k * step]);
        float xr = re[i + k + half] * wr - im[i + k + half] * wi;
        float xi = re[i + k + half] * wi + im[i + k + half] * wr;
        re[i + k + half] = re[i + k] - xr; im[i + k + half] = im[i + k] - xi;
        re[i + k] += xr; im[i + k] += xi;
      }
  }
}

static float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

ORC_API int orc_mfcc(const orc_mfcc_opts *o, const int16_t *pcm, int64_t n, float *out /* [T][num_ceps] */) {
  const int N = win_size(o), shift = win_shift(o), NP = round_up_pow2(N), NB = NP / 2;
  const int nbins = o->num_mel_bins, nceps = o->num_ceps;
  const int64_t T = orc_mfcc_num_frames(o, n);
  if (T == 0) return 0;
  float *window = malloc(sizeof(float) * N);
  for (int i = 0; i < N; i++) {
    double a = 2.0 * M_PI / (N - 1);
    window[i] = (float)pow(0.5 - 0.5 * cos(a * (double)i), 0.85); /* povey */
  }
  /* mel banks (float arithmetic as in MelBanks::MelBanks) */
  float nyquist = 0.5f * o->sample_frequency;
  float high = o->high_freq > 0.0f ? o->high_freq : nyquist + o->high_freq;
  float fft_bin_width = o->sample_frequency / NP;
  float mel_low = mel_scale(o->low_freq), mel_high = mel_scale(high);
  float mel_delta = (mel_high - mel_low) / (nbins + 1);
  float *melw = calloc((size_t)nbins * NB, sizeof(float));
  int *mfirst = malloc(sizeof(int) * nbins), *mlast = malloc(sizeof(int) * nbins);
  for (int b = 0; b < nbins; b++) {
    float left = mel_low + b * mel_delta, center = mel_low + (b + 1) * mel_delta, right = mel_low + (b + 2) * mel_delta;
    mfirst[b] = -1; mlast[b] = -1;
    for (int i = 0; i < NB; i++) {
      float mel = mel_scale(fft_bin_width * i);
      if (mel > left && mel < right) {
        float w = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
        melw[b * NB + i] = w;
        if (mfirst[b] < 0) mfirst[b] = i;
        mlast[b] = i;
      }
    }
  }
  /* DCT-II rows 0..nceps-1 and lifter */
  float *dct = malloc(sizeof(float) * nceps * nbins), *lift = malloc(sizeof(float) * nceps);
  for (int k = 0; k < nceps; k++)
    for (int j = 0; j < nbins; j++) {
      if (k == 0) dct[j] = (float)sqrt(1.0 / nbins);
      else dct[k * nbins + j] = (float)(sqrt(2.0 / nbins) * cos(M_PI / nbins * (j + 0.5) * k));
    }
  for (int k = 0; k < nceps; k++)
    lift[k] = (o->cepstral_lifter != 0.0f) ? (float)(1.0 + 0.5 * o->cepstral_lifter * sin(M_PI * k / o->cepstral_lifter)) : 1.0f;
  /* twiddles for the NP/2-point complex FFT + real-FFT post-processing */
  double *cs = malloc(sizeof(double) * NP), *sn = malloc(sizeof(double) * NP);
  for (int i = 0; i < NP; i++) { cs[i] = cos(2.0 * M_PI * i / NP); sn[i] = sin(2.0 * M_PI * i / NP); }
  double *cs2 = malloc(sizeof(double) * NB), *sn2 = malloc(sizeof(double) * NB);
  for (int i = 0; i < NB; i++) { cs2[i] = cos(2.0 * M_PI * i / NB); sn2[i] = sin(2.0 * M_PI * i / NB); }
  float *x = malloc(sizeof(float) * NP), *re = malloc(sizeof(float) * NB), *im = malloc(sizeof(float) * NB);
  float *pw = malloc(sizeof(float) * (NB + 1)), *mel = malloc(sizeof(float) * nbins);
  for (int64_t f = 0; f < T; f++) {
    int64_t start = o->snip_edges ? f * shift : f * shift + shift / 2 - N / 2;
    for (int s = 0; s < N; s++) {
      int64_t k = start + s;
      while (k < 0 || k >= n) k = (k < 0) ? -k - 1 : 2 * n - 1 - k;
      x[s] = (float)pcm[k];
    }
    if (o->remove_dc_offset) {
      double sum = 0.0; for (int s = 0; s < N; s++) sum += x[s];
      float m = (float)(sum / N); for (int s = 0; s < N; s++) x[s] -= m;
    }
    float log_energy = 0.0f;
    if (o->use_energy && o->raw_energy) {
      float e = 0.0f; for (int s = 0; s < N; s++) e += x[s] * x[s];
      log_energy = logf(e > FLT_EPSILON ? e : FLT_EPSILON);
    }
    if (o->preemph_coeff != 0.0f) {
      for (int s = N - 1; s > 0; s--) x[s] -= o->preemph_coeff * x[s - 1];
      x[0] -= o->preemph_coeff * x[0];
    }
    for (int s = 0; s < N; s++) x[s] *= window[s];
    if (o->use_energy && !o->raw_energy) {
      float e = 0.0f; for (int s = 0; s < N; s++) e += x[s] * x[s];
      log_energy = logf(e > FLT_EPSILON ? e : FLT_EPSILON);
    }
    for (int s = N; s < NP; s++) x[s] = 0.0f;
    /* real FFT of NP points via NB-point complex FFT */
    for (int i = 0; i < NB; i++) { re[i] = x[2 * i]; im[i] = x[2 * i + 1]; }
    cfft(re, im, NB, cs2, sn2);
    pw[0] = (re[0] + im[0]) * (re[0] + im[0]);
    pw[NB] = (re[0] - im[0]) * (re[0] - im[0]);
    for (int k = 1; k < NB; k++) {
      float ar = re[k], ai = im[k], br = re[NB - k], bi = -im[NB - k];
      float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);  /* even part */
      float dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);  /* (odd part) * i */
      float wr = (float)cs[k], wi = (float)(-sn[k]);
      /* odd = -i * d ; X = even + w * odd */
      float orr = di, oi = -dr;
      float xr = er + (orr * wr - oi * wi), xi = ei + (orr * wi + oi * wr);
      pw[k] = xr * xr + xi * xi;
    }
    for (int b = 0; b < nbins; b++) {
      float e = 0.0f;
      for (int i = mfirst[b]; i <= mlast[b] && i >= 0; i++) e += melw[b * NB + i] * pw[i];
      if (e < FLT_EPSILON) e = FLT_EPSILON;
      mel[b] = logf(e);
    }
    float *row = out + f * nceps;
    for (int k = 0; k < nceps; k++) {
      float acc = 0.0f;
      for (int j = 0; j < nbins; j++) acc += dct[k * nbins + j] * mel[j];
      row[k] = acc * lift[k];
    }
    if (o->use_energy) {
      if (o->energy_floor > 0.0f && log_energy < logf(o->energy_floor)) log_energy = logf(o->energy_floor);
      row[0] = log_energy;
    }
  }
  free(window); free(melw); free(mfirst); free(mlast); free(dct); free(lift); free(cs); free(sn); free(cs2); free(sn2);
  free(x); free(re); free(im); free(pw); free(mel);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * CMVN.  Call sites: corpus/acoustic_corpus.py:1336 (CmvnComputer.export_cmvn),
 * command_line/align_one.py:168,183.  Restates Kaldi transform/cmvn.cc AccCmvnStats / ApplyCmvn
 * with norm_vars=false (alignment/multiprocessing.py:1288).  SURVEY.md A.3.
 * stats layout: [2][dim+1] doubles, stats[0][dim] = count.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_cmvn_acc(const float *feats, int64_t T, int dim, double *stats) {
  for (int64_t t = 0; t < T; t++) {
    const float *r = feats + t * dim;
    for (int d = 0; d < dim; d++) { stats[d] += r[d]; stats[(dim + 1) + d] += (double)r[d] * r[d]; }
  }
  stats[dim] += (double)T;
}

ORC_API void orc_cmvn_apply(float *feats, int64_t T, int dim, const double *stats) {
  double count = stats[dim];
  for (int d = 0; d < dim; d++) {
    float offset = (float)(-(stats[d] / count));
    for (int64_t t = 0; t < T; t++) feats[t * dim + d] += offset;
  }
}

/* ------------------------------------------------------------------------------------------
 * Deltas / splice / linear transforms.  Call site: db.py:2101-2136 (Job.construct_feature_archive),
 * op order restated in-tree at alignment/multiprocessing.py:1287-1304.  Restates Kaldi
 * feat/feature-functions.cc (DeltaFeatures::Process order=2 window=2, SpliceFrames) and
 * transform/transform-common (ApplyAffineTransform).  SURVEY.md A.3.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_add_deltas(const float *in, int64_t T, int dim, float *out /* [T][3*dim] */) {
  static const float s1[5] = {-0.2f, -0.1f, 0.0f, 0.1f, 0.2f};
  float s2[9];
  for (int i = 0; i < 9; i++) s2[i] = 0.0f;
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) s2[i + j] += s1[i] * s1[j];
  for (int64_t t = 0; t < T; t++) {
    float *o = out + t * 3 * dim;
    for (int d = 0; d < dim; d++) o[d] = in[t * dim + d];
    for (int d = 0; d < dim; d++) { o[dim + d] = 0.0f; o[2 * dim + d] = 0.0f; }
    for (int j = -2; j <= 2; j++) {
      float sc = s1[j + 2]; if (sc == 0.0f) continue;
      int64_t tt = t + j; if (tt < 0) tt = 0; if (tt >= T) tt = T - 1;
      for (int d = 0; d < dim; d++) o[dim + d] += sc * in[tt * dim + d];
    }
    for (int j = -4; j <= 4; j++) {
      float sc = s2[j + 4]; if (sc == 0.0f) continue;
      int64_t tt = t + j; if (tt < 0) tt = 0; if (tt >= T) tt = T - 1;
      for (int d = 0; d < dim; d++) o[2 * dim + d] += sc * in[tt * dim + d];
    }
  }
}

ORC_API void orc_splice(const float *in, int64_t T, int dim, int left, int right, float *out) {
  int w = left + right + 1;
  for (int64_t t = 0; t < T; t++)
    for (int j = 0; j < w; j++) {
      int64_t tt = t + j - left; if (tt < 0) tt = 0; if (tt >= T) tt = T - 1;
      memcpy(out + (t * w + j) * dim, in + tt * dim, sizeof(float) * dim);
    }
}

/* y = M x (cols == in_dim) or y = M[:, :in_dim] x + M[:, in_dim] (cols == in_dim + 1) */
ORC_API void orc_transform(const float *in, int64_t T, int in_dim, const float *M, int rows, int cols, float *out) {
  for (int64_t t = 0; t < T; t++)
    for (int r = 0; r < rows; r++) {
      float acc = 0.0f;
      for (int c = 0; c < in_dim; c++) acc += M[r * cols + c] * in[t * in_dim + c];
      if (cols == in_dim + 1) acc += M[r * cols + in_dim];
      out[t * rows + r] = acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * Diagonal-GMM log-likelihood.  Call sites: inside GmmAligner (alignment/multiprocessing.py:846)
 * and gmm_compute_likes (alignment/multiprocessing.py:1415).  Restates Kaldi
 * gmm/decodable-am-diag-gmm.cc LogLikelihoodZeroBased + matrix/kaldi-vector.cc LogSumExp.  A.4.
 * Model layout: Gaussians of pdf j are rows off[j]..off[j+1]-1.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t dim, num_pdfs;
  const int32_t *off;
  const float *gconsts, *means_invvars, *inv_vars;
} orc_gmm;

static const float kMinLogDiffFloat = -15.9423847198486328125f; /* logf(FLT_EPSILON) */

static float gmm_loglike_pdf(const orc_gmm *g, const float *x, const float *x2, int pdf, float *comp /* optional */) {
  int a = g->off[pdf], b = g->off[pdf + 1], D = g->dim;
  float mx = -INFINITY;
  float buf[4096];
  float *ll = comp ? comp : buf;
  for (int m = a; m < b; m++) {
    const float *miv = g->means_invvars + (size_t)m * D, *iv = g->inv_vars + (size_t)m * D;
    float d1 = 0.0f, d2 = 0.0f;
    for (int d = 0; d < D; d++) d1 += miv[d] * x[d];
    for (int d = 0; d < D; d++) d2 += iv[d] * x2[d];
    float v = g->gconsts[m] + d1;
    v = v + (-0.5f) * d2;
    ll[m - a] = v;
    if (v > mx) mx = v;
  }
  float cutoff = mx + kMinLogDiffFloat;
  double sum = 0.0;
  for (int m = 0; m < b - a; m++) if (ll[m] >= cutoff) sum += expf(ll[m] - mx);
  return (float)(mx + log(sum));
}

ORC_API void orc_gmm_loglikes(const orc_gmm *g, const float *feats, int64_t T, float *out /* [T][num_pdfs] */) {
  float *x2 = malloc(sizeof(float) * g->dim);
  for (int64_t t = 0; t < T; t++) {
    const float *x = feats + t * g->dim;
    for (int d = 0; d < g->dim; d++) x2[d] = x[d] * x[d];
    for (int p = 0; p < g->num_pdfs; p++) out[t * g->num_pdfs + p] = gmm_loglike_pdf(g, x, x2, p, NULL);
  }
  free(x2);
}

/* ------------------------------------------------------------------------------------------
 * Viterbi alignment.  Call sites: alignment/multiprocessing.py:846-853 (AlignFunction ->
 * GmmAligner.export_alignments), online/alignment.py:97-107.  Restates Kaldi
 * decoder/faster-decoder.cc (FasterDecoder, incl. util/hash-list-inl.h iteration order),
 * decoder/decoder-wrappers.cc (AlignUtteranceWrapper: beam, then retry_beam), hmm/hmm-utils.cc
 * (AddTransitionProbs -- the caller passes tid_cost[tid] = -scaled log prob).  SURVEY.md A.5-A.7.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t num_states, start;
  const int32_t *arc_off; /* [num_states+1], arcs grouped by source state in OpenFst order */
  const int32_t *ilabel, *olabel, *nextstate;
  const float *weight;    /* graph weights BEFORE AddTransitionProbs */
  const float *final;     /* +inf = non-final */
} orc_fst;

typedef struct Tok { double cost; int32_t arc; struct Tok *prev; } Tok;
typedef struct Elem { int32_t key; Tok *val; struct Elem *tail; } Elem;
typedef struct { Elem *last_elem; size_t prev_bucket; } Bucket;

typedef struct {
  Elem *list_head; size_t bucket_list_tail; size_t hash_size; Bucket *buckets; size_t buckets_cap;
  Elem *freed;
  /* arenas */
  void **blocks; int nblocks, capblocks; char *cur; size_t left;
} HashList;

static void *arena_alloc(HashList *h, size_t sz) {
  if (h->left < sz) {
    size_t bs = 1 << 20;
    if (h->nblocks == h->capblocks) { h->capblocks = h->capblocks ? 2 * h->capblocks : 16; h->blocks = realloc(h->blocks, sizeof(void *) * h->capblocks); }
    h->cur = malloc(bs); h->blocks[h->nblocks++] = h->cur; h->left = bs;
  }
  void *p = h->cur; h->cur += sz; h->left -= sz; return p;
}
static void hl_set_size(HashList *h, size_t sz) {
  h->hash_size = sz;
  if (sz > h->buckets_cap) {
    h->buckets = realloc(h->buckets, sizeof(Bucket) * sz);
    for (size_t i = h->buckets_cap; i < sz; i++) { h->buckets[i].last_elem = NULL; h->buckets[i].prev_bucket = (size_t)-1; }
    h->buckets_cap = sz;
  }
}
static Elem *hl_clear(HashList *h) {
  for (size_t cur = h->bucket_list_tail; cur != (size_t)-1; cur = h->buckets[cur].prev_bucket) h->buckets[cur].last_elem = NULL;
  h->bucket_list_tail = (size_t)-1;
  Elem *ans = h->list_head; h->list_head = NULL; return ans;
}
static void hl_delete(HashList *h, Elem *e) { e->tail = h->freed; h->freed = e; }
static Elem *hl_find_or_insert(HashList *h, int32_t key, Tok *val) {
  size_t index = (size_t)key % h->hash_size;
  Bucket *b = &h->buckets[index];
  if (b->last_elem != NULL) {
    Elem *head = (b->prev_bucket == (size_t)-1) ? h->list_head : h->buckets[b->prev_bucket].last_elem->tail;
    Elem *tail = b->last_elem->tail;
    for (Elem *e = head; e != tail; e = e->tail) if (e->key == key) return e;
  }
  Elem *elem;
  if (h->freed) { elem = h->freed; h->freed = elem->tail; } else elem = arena_alloc(h, sizeof(Elem));
  elem->key = key; elem->val = val;
  if (b->last_elem == NULL) {
    if (h->bucket_list_tail == (size_t)-1) h->list_head = elem;
    else h->buckets[h->bucket_list_tail].last_elem->tail = elem;
    elem->tail = NULL; b->last_elem = elem; b->prev_bucket = h->bucket_list_tail; h->bucket_list_tail = index;
  } else {
    elem->tail = b->last_elem->tail; b->last_elem->tail = elem; b->last_elem = elem;
  }
  return elem;
}

typedef struct {
  const orc_fst *fst; const float *tid_cost; /* -scaled transition log prob per tid */
  const orc_gmm *gmm; const int32_t *tid2pdf; const float *feats; int64_t T; float acoustic_scale;
  /* lazy per-frame pdf cache (DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased) */
  float *cache; int64_t *hit; float *x2; int64_t x2_frame;
  const float *dense; /* optional precomputed [T][num_pdfs] loglikes instead of gmm+feats */
  int num_pdfs;
  HashList hl;
  double *tmp; size_t tmp_cap;
  Elem **queue; size_t qn, qcap;
  float beam; int min_active, max_active; float beam_delta, hash_ratio;
  int64_t frames_decoded;
} Decoder;

static float dec_loglike_raw(Decoder *d, int64_t frame, int tid) {
  int pdf = d->tid2pdf[tid];
  if (d->dense) return d->dense[frame * d->num_pdfs + pdf];
  if (d->hit[pdf] == frame) return d->cache[pdf];
  const float *x = d->feats + frame * d->gmm->dim;
  if (d->x2_frame != frame) { for (int k = 0; k < d->gmm->dim; k++) d->x2[k] = x[k] * x[k]; d->x2_frame = frame; }
  float v = gmm_loglike_pdf(d->gmm, x, d->x2, pdf, NULL);
  d->cache[pdf] = v; d->hit[pdf] = frame; return v;
}
static float dec_loglike(Decoder *d, int64_t frame, int tid) { return d->acoustic_scale * dec_loglike_raw(d, frame, tid); }

static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

static double get_cutoff(Decoder *d, Elem *list, size_t *tok_count, float *adaptive_beam, Elem **best_elem) {
  double best = INFINITY; size_t count = 0; size_t n = 0;
  for (Elem *e = list; e; e = e->tail, count++) {
    double w = e->val->cost;
    if (n == d->tmp_cap) { d->tmp_cap = d->tmp_cap ? 2 * d->tmp_cap : 1024; d->tmp = realloc(d->tmp, sizeof(double) * d->tmp_cap); }
    d->tmp[n++] = w;
    if (w < best) { best = w; *best_elem = e; }
  }
  *tok_count = count;
  double beam_cutoff = best + d->beam, min_active_cutoff = INFINITY, max_active_cutoff = INFINITY;
  if (n > (size_t)d->max_active) { qsort(d->tmp, n, sizeof(double), cmp_double); max_active_cutoff = d->tmp[d->max_active]; }
  if (max_active_cutoff < beam_cutoff) { *adaptive_beam = (float)(max_active_cutoff - best + d->beam_delta); return max_active_cutoff; }
  if (n > (size_t)d->min_active) {
    if (d->min_active == 0) min_active_cutoff = best;
    else { qsort(d->tmp, n, sizeof(double), cmp_double); min_active_cutoff = d->tmp[d->min_active]; } /* nth_element */
  }
  if (min_active_cutoff > beam_cutoff) { *adaptive_beam = (float)(min_active_cutoff - best + d->beam_delta); return min_active_cutoff; }
  *adaptive_beam = d->beam; return beam_cutoff;
}

static Tok *new_tok(Decoder *d, int arc, double cost, Tok *prev) {
  Tok *t = arena_alloc(&d->hl, sizeof(Tok)); t->cost = cost; t->arc = arc; t->prev = prev; return t;
}

static void queue_push(Decoder *d, Elem *e) {
  if (d->qn == d->qcap) { d->qcap = d->qcap ? 2 * d->qcap : 1024; d->queue = realloc(d->queue, sizeof(Elem *) * d->qcap); }
  d->queue[d->qn++] = e;
}

static void process_nonemitting(Decoder *d, double cutoff) {
  const orc_fst *f = d->fst;
  d->qn = 0;
  for (Elem *e = d->hl.list_head; e; e = e->tail) queue_push(d, e);
  while (d->qn) {
    Elem *e = d->queue[--d->qn];
    int st = e->key; Tok *tok = e->val;
    if (tok->cost > cutoff) continue;
    for (int a = f->arc_off[st]; a < f->arc_off[st + 1]; a++) {
      if (f->ilabel[a] != 0) continue;
      double c = tok->cost + f->weight[a];
      if (c > cutoff) continue;
      Tok *nt = new_tok(d, a, c, tok);
      Elem *ef = hl_find_or_insert(&d->hl, f->nextstate[a], nt);
      if (ef->val == nt) queue_push(d, ef);
      else if (ef->val->cost > nt->cost) { ef->val = nt; queue_push(d, ef); }
    }
  }
}

static double process_emitting(Decoder *d) {
  const orc_fst *f = d->fst;
  int64_t frame = d->frames_decoded;
  Elem *last = hl_clear(&d->hl);
  size_t tok_cnt; float adaptive_beam; Elem *best_elem = NULL;
  double weight_cutoff = get_cutoff(d, last, &tok_cnt, &adaptive_beam, &best_elem);
  size_t new_sz = (size_t)((float)tok_cnt * d->hash_ratio);
  if (new_sz > d->hl.hash_size) hl_set_size(&d->hl, new_sz);
  double next_cutoff = INFINITY;
  if (best_elem) {
    int st = best_elem->key; Tok *tok = best_elem->val;
    for (int a = f->arc_off[st]; a < f->arc_off[st + 1]; a++) {
      if (f->ilabel[a] == 0) continue;
      float ac = -dec_loglike(d, frame, f->ilabel[a]);
      float gw = f->weight[a] + d->tid_cost[f->ilabel[a]];
      double nw = gw + tok->cost + ac;
      if (nw + adaptive_beam < next_cutoff) next_cutoff = nw + adaptive_beam;
    }
  }
  for (Elem *e = last, *et; e; e = et) {
    int st = e->key; Tok *tok = e->val;
    if (tok->cost < weight_cutoff) {
      for (int a = f->arc_off[st]; a < f->arc_off[st + 1]; a++) {
        if (f->ilabel[a] == 0) continue;
        float ac = -dec_loglike(d, frame, f->ilabel[a]);
        float gw = f->weight[a] + d->tid_cost[f->ilabel[a]];
        double nw = gw + tok->cost + ac;
        if (nw < next_cutoff) {
          Tok *nt = new_tok(d, a, tok->cost + gw + ac, tok);
          Elem *ef = hl_find_or_insert(&d->hl, f->nextstate[a], nt);
          if (nw + adaptive_beam < next_cutoff) next_cutoff = nw + adaptive_beam;
          if (ef->val != nt && ef->val->cost > nt->cost) ef->val = nt;
        }
      }
    }
    et = e->tail; hl_delete(&d->hl, e);
  }
  d->frames_decoded++;
  return next_cutoff;
}

static int reached_final(Decoder *d) {
  for (Elem *e = d->hl.list_head; e; e = e->tail)
    if (e->val->cost != INFINITY && d->fst->final[e->key] != INFINITY) return 1;
  return 0;
}

static void decode(Decoder *d) {
  HashList *h = &d->hl;
  /* ClearToks + fresh arenas */
  for (int i = 0; i < h->nblocks; i++) free(h->blocks[i]);
  h->nblocks = 0; h->left = 0; h->freed = NULL; h->list_head = NULL;
  for (size_t i = 0; i < h->buckets_cap; i++) { h->buckets[i].last_elem = NULL; h->buckets[i].prev_bucket = (size_t)-1; }
  h->bucket_list_tail = (size_t)-1;
  if (h->hash_size == 0) hl_set_size(h, 1000);
  for (int p = 0; d->hit && p < d->num_pdfs; p++) d->hit[p] = -1;
  d->x2_frame = -1;
  d->frames_decoded = 0;
  hl_find_or_insert(h, d->fst->start, new_tok(d, -1, 0.0, NULL));
  process_nonemitting(d, INFINITY);
  while (d->frames_decoded < d->T) { double c = process_emitting(d); process_nonemitting(d, c); }
}

/* returns 0 ok, 1 retried ok, 2 failed (no final state reached with either beam), 3 empty graph, 4 zero frames */
ORC_API int orc_align(const orc_fst *fst, const float *tid_cost, const orc_gmm *gmm, const int32_t *tid2pdf,
                      const float *feats, const float *dense_loglikes, int64_t T, float acoustic_scale, float beam,
                      float retry_beam, int32_t *ali /* [T] */, int32_t *words, int32_t *num_words, int32_t max_words,
                      float *per_frame /* [T] */, float *total_like) {
  *num_words = 0; *total_like = 0.0f;
  if (fst->start < 0 || fst->num_states == 0) return 3;
  if (T == 0) return 4;
  Decoder d; memset(&d, 0, sizeof(d));
  d.fst = fst; d.tid_cost = tid_cost; d.gmm = gmm; d.tid2pdf = tid2pdf; d.feats = feats; d.T = T; d.acoustic_scale = acoustic_scale;
  d.dense = dense_loglikes; d.num_pdfs = gmm->num_pdfs;
  if (!d.dense) { d.cache = malloc(sizeof(float) * gmm->num_pdfs); d.hit = malloc(sizeof(int64_t) * gmm->num_pdfs); d.x2 = malloc(sizeof(float) * gmm->dim); }
  d.beam = beam; d.min_active = 20; d.max_active = 2147483647; d.beam_delta = 0.5f; d.hash_ratio = 2.0f;
  int status = 0;
  decode(&d);
  int ok = reached_final(&d);
  if (!ok && retry_beam != 0.0f) { d.beam = retry_beam; decode(&d); ok = reached_final(&d); status = 1; }
  if (!ok) status = 2;
  else {
    /* GetBestPath */
    Tok *best = NULL; double best_cost = INFINITY; int best_state = -1;
    for (Elem *e = d.hl.list_head; e; e = e->tail) {
      double c = e->val->cost + fst->final[e->key];
      if (c < best_cost && c != INFINITY) { best_cost = c; best = e->val; best_state = e->key; }
    }
    /* walk back collecting arcs */
    int64_t n = 0; for (Tok *t = best; t && t->arc >= 0; t = t->prev) n++;
    int32_t *arcs = malloc(sizeof(int32_t) * (n ? n : 1)); float *gcost = malloc(sizeof(float) * (n ? n : 1)), *acost = malloc(sizeof(float) * (n ? n : 1));
    int64_t i = n;
    for (Tok *t = best; t && t->arc >= 0; t = t->prev) {
      --i; arcs[i] = t->arc;
      float tot = (float)(t->cost - (t->prev ? t->prev->cost : 0.0));
      int il = fst->ilabel[t->arc];
      float g = fst->weight[t->arc] + (il ? tid_cost[il] : 0.0f);
      gcost[i] = g; acost[i] = tot - g;
    }
    /* GetLinearSymbolSequence: LatticeWeight Times in float, start -> end, then final */
    float v1 = 0.0f, v2 = 0.0f; int64_t tf = 0;
    for (i = 0; i < n; i++) {
      int a = arcs[i];
      if (fst->ilabel[a] != 0) {
        if (tf < T) { ali[tf] = fst->ilabel[a]; per_frame[tf] = -acost[i] / acoustic_scale; }
        tf++;
      }
      if (fst->olabel[a] != 0) { if (*num_words < max_words) words[*num_words] = fst->olabel[a]; (*num_words)++; }
      v1 += gcost[i]; v2 += acost[i];
    }
    v1 += fst->final[best_state];
    *total_like = -(v1 + v2) / acoustic_scale;
    free(arcs); free(gcost); free(acost);
    if (tf != T) status = 2;
  }
  for (int b = 0; b < d.hl.nblocks; b++) free(d.hl.blocks[b]);
  free(d.hl.blocks); free(d.hl.buckets); free(d.tmp); free(d.queue); free(d.cache); free(d.hit); free(d.x2);
  return status;
}

/* ------------------------------------------------------------------------------------------
 * GMM accumulator statistics.  Call site: alignment/multiprocessing.py:652-666 (AccStatsFunction ->
 * GmmStatsAccumulator.accumulate_stats).  Restates Kaldi gmmbin/gmm-acc-stats-ali.cc,
 * gmm/mle-am-diag-gmm.cc AccumulateForGmm, gmm/mle-diag-gmm.cc AccumulateFromDiag /
 * AccumulateFromPosteriors, gmm/diag-gmm.cc ComponentPosteriors, ApplySoftMax.  SURVEY.md A.8.
 * occ[G], mean_acc[G][D], var_acc[G][D], trans_acc[num_tids+1] doubles; returns tot_like via ptr.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_acc_stats(const orc_gmm *g, const int32_t *tid2pdf, const float *feats, const int32_t *ali, int64_t T,
                           double *occ, double *mean_acc, double *var_acc, double *trans_acc, double *tot_like) {
  int D = g->dim;
  float *x2 = malloc(sizeof(float) * D); float comp[4096];
  double like = 0.0;
  for (int64_t t = 0; t < T; t++) {
    int tid = ali[t], pdf = tid2pdf[tid];
    trans_acc[tid] += 1.0;
    const float *x = feats + t * D;
    for (int d = 0; d < D; d++) x2[d] = x[d] * x[d];
    int a = g->off[pdf], b = g->off[pdf + 1];
    float mx = -INFINITY;
    for (int m = a; m < b; m++) {
      const float *miv = g->means_invvars + (size_t)m * D, *iv = g->inv_vars + (size_t)m * D;
      float d1 = 0.0f, d2 = 0.0f;
      for (int d = 0; d < D; d++) d1 += miv[d] * x[d];
      for (int d = 0; d < D; d++) d2 += iv[d] * x2[d];
      float v = g->gconsts[m] + d1; v = v + (-0.5f) * d2;
      comp[m - a] = v; if (v > mx) mx = v;
    }
    float sum = 0.0f;
    for (int m = 0; m < b - a; m++) { comp[m] = expf(comp[m] - mx); sum += comp[m]; }
    float inv = 1.0f / sum;
    for (int m = 0; m < b - a; m++) comp[m] *= inv;
    float log_sum = mx + logf(sum);
    like += log_sum;
    for (int m = a; m < b; m++) {
      double p = comp[m - a];
      occ[m] += p;
      for (int d = 0; d < D; d++) { double xd = x[d]; mean_acc[(size_t)m * D + d] += p * xd; var_acc[(size_t)m * D + d] += p * (xd * xd); }
    }
  }
  *tot_like += like;
  free(x2);
}

/* ------------------------------------------------------------------------------------------
 * Equal alignment.  Call site: acoustic_modeling/monophone.py:108 (MonoAlignEqualFunction ->
 * kalpy gmm_align_equal).  Restates Kaldi fstext/fstext-utils-inl.h EqualAlign (random self-loop-free
 * path drawn with srand(seed)/rand() through kaldi::RandInt, retried while it has more input labels
 * than frames; extra frames spread over the path's self-loops, the first (extra % loops) of them get
 * one more) followed by GetLinearSymbolSequence.  Uses libc rand() directly (single-threaded).
 * Returns 0 ok, 2 failed, 3 empty graph, 4 zero frames.
 * ---------------------------------------------------------------------------------------- */
static int kaldi_rand_int(int lo, int hi) { return lo == hi ? lo : lo + rand() % (hi - lo + 1); }

ORC_API int orc_equal_align(const orc_fst *fst, int64_t T, unsigned seed, int num_retries, int32_t *ali, int32_t *words,
                            int32_t *num_words, int32_t max_words) {
  *num_words = 0;
  if (T <= 0) return 4;
  if (fst->start < 0 || fst->num_states == 0) return 3;
  srand(seed);
  int cap = 1024, np = 0; int64_t n_il = 0;
  int32_t *path = malloc(sizeof(int32_t) * cap), *taken = malloc(sizeof(int32_t) * cap);
  int retry = 0;
  do {
    n_il = 0; np = 0; path[np++] = fst->start;
    for (;;) {
      int s = path[np - 1];
      int na = fst->arc_off[s + 1] - fst->arc_off[s];
      int tot = na + (fst->final[s] < INFINITY ? 1 : 0);
      if (tot == 0) { free(path); free(taken); return 2; }
      int off = kaldi_rand_int(0, tot - 1);
      if (off >= na) break;
      int a = fst->arc_off[s] + off;
      if (fst->nextstate[a] == s) continue;
      if (np == cap) { cap *= 2; path = realloc(path, sizeof(int32_t) * cap); taken = realloc(taken, sizeof(int32_t) * cap); }
      taken[np - 1] = a; path[np++] = fst->nextstate[a];
      if (fst->ilabel[a] != 0) n_il++;
    }
  } while (++retry < num_retries && n_il > T);
  int status = 0;
  if (n_il > T) status = 2;
  int32_t *loop = malloc(sizeof(int32_t) * np);
  int64_t n_loops = 0;
  for (int i = 0; i < np && !status; i++) {
    loop[i] = -1;
    for (int a = fst->arc_off[path[i]]; a < fst->arc_off[path[i] + 1]; a++)
      if (fst->nextstate[a] == path[i] && fst->ilabel[a] != 0) { loop[i] = a; n_loops++; break; }
  }
  if (!status && n_loops == 0 && n_il < T) status = 2;
  if (!status) {
    int64_t extra = T - n_il, min_loops = extra ? extra / n_loops : 0, one_more = extra - min_loops * n_loops, counter = 0, t = 0;
    for (int i = 0; i < np; i++) {
      if (loop[i] >= 0) {
        int64_t k = min_loops + (counter < one_more ? 1 : 0);
        counter++;
        for (int64_t j = 0; j < k; j++) {
          ali[t++] = fst->ilabel[loop[i]];
          if (fst->olabel[loop[i]] != 0) { if (*num_words < max_words) words[*num_words] = fst->olabel[loop[i]]; (*num_words)++; }
        }
      }
      if (i + 1 < np) {
        int a = taken[i];
        if (fst->ilabel[a] != 0) ali[t++] = fst->ilabel[a];
        if (fst->olabel[a] != 0) { if (*num_words < max_words) words[*num_words] = fst->olabel[a]; (*num_words)++; }
      }
    }
    if (t != T) status = 2;
  }
  free(path); free(taken); free(loop);
  return status;
}

/* ------------------------------------------------------------------------------------------
 * fMLLR estimation.  Call site: corpus/features.py:460-548 (CalcFmllrFunction -> kalpy FmllrComputer.export_transforms),
 * options corpus/features.py:759-766.  Restates Kaldi gmmbin/gmm-est-fmllr(-gpost).cc, transform/fmllr-diag-gmm.cc
 * (FmllrDiagGmmAccs::AccumulateFromPosteriors / CommitSingleFrameStats, ComputeFmllrMatrixDiagGmmFull, FmllrInnerUpdate,
 * FmllrAuxFuncDiagGmm) and hmm/posterior.cc WeightSilencePost.
 *   g_post: model the component posteriors come from (the alignment model when two models are used), g: model whose
 *   means / variances enter the statistics; both evaluated on the SAME features.  tid_weight[tid]: frame weight
 *   (silence_weight for silence phones, else 1; weight 0 drops the frame).
 *   stats (doubles): beta | K[D][D+1] | G[D][(D+1)(D+2)/2]  (G_d packed lower triangle, row-major: Kaldi SpMatrix).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_fmllr_acc(const orc_gmm *g_post, const orc_gmm *g, const int32_t *tid2pdf, const float *tid_weight,
                           const float *feats, const int32_t *ali, int64_t T, double *stats) {
  const int D = g->dim, D1 = D + 1, NP = D1 * (D1 + 1) / 2;
  double *beta = stats, *K = stats + 1, *G = K + (size_t)D * D1;
  float comp[4096];
  float *a = malloc(sizeof(float) * D), *b = malloc(sizeof(float) * D), *x2 = malloc(sizeof(float) * D);
  double *xp = malloc(sizeof(double) * D1);
  for (int64_t t = 0; t < T; t++) {
    const int tid = ali[t];
    const float w = tid_weight ? tid_weight[tid] : 1.0f;
    if (w == 0.0f) continue;
    const int pdf = tid2pdf[tid];
    const float *x = feats + t * D;
    for (int d = 0; d < D; d++) x2[d] = x[d] * x[d];
    const int m0 = g_post->off[pdf], m1 = g_post->off[pdf + 1];
    float mx = -INFINITY;
    for (int m = m0; m < m1; m++) {
      const float *miv = g_post->means_invvars + (size_t)m * D, *iv = g_post->inv_vars + (size_t)m * D;
      float d1 = 0.0f, d2 = 0.0f;
      for (int d = 0; d < D; d++) d1 += miv[d] * x[d];
      for (int d = 0; d < D; d++) d2 += iv[d] * x2[d];
      float v = g_post->gconsts[m] + d1; v = v + (-0.5f) * d2;
      comp[m - m0] = v; if (v > mx) mx = v;
    }
    float sum = 0.0f;
    for (int m = 0; m < m1 - m0; m++) { comp[m] = expf(comp[m] - mx); sum += comp[m]; }
    const float inv = 1.0f / sum;
    double count = 0.0;
    for (int d = 0; d < D; d++) { a[d] = 0.0f; b[d] = 0.0f; }
    for (int m = 0; m < m1 - m0; m++) {
      const float p = comp[m] * inv * w;
      count += p;
      const float *miv = g->means_invvars + (size_t)(m0 + m) * D, *iv = g->inv_vars + (size_t)(m0 + m) * D;
      for (int d = 0; d < D; d++) { a[d] += miv[d] * p; b[d] += iv[d] * p; }
    }
    for (int d = 0; d < D; d++) xp[d] = x[d];
    xp[D] = 1.0;
    *beta += count;
    for (int d = 0; d < D; d++) for (int j = 0; j < D1; j++) K[(size_t)d * D1 + j] += (double)a[d] * xp[j];
    for (int d = 0; d < D; d++) {
      if (b[d] == 0.0f) continue;
      double *Gd = G + (size_t)d * NP;
      const double bd = b[d];
      int k = 0;
      for (int i = 0; i < D1; i++) for (int j = 0; j <= i; j++, k++) Gd[k] += bd * (xp[i] * xp[j]);
    }
  }
  free(a); free(b); free(x2); free(xp);
}

/* in-place inverse of an n x n matrix (Gauss-Jordan, partial pivoting); returns log|det| through *logdet (may be NULL); 0 ok */
static int mat_invert(double *A, int n, double *logdet) {
  int *piv = malloc(sizeof(int) * n);
  double ld = 0.0; int ok = 0;
  for (int c = 0; c < n; c++) {
    int p = c; double best = fabs(A[c * n + c]);
    for (int r = c + 1; r < n; r++) if (fabs(A[r * n + c]) > best) { best = fabs(A[r * n + c]); p = r; }
    if (best == 0.0) { ok = -1; break; }
    piv[c] = p;
    if (p != c) for (int j = 0; j < n; j++) { double tmp = A[c * n + j]; A[c * n + j] = A[p * n + j]; A[p * n + j] = tmp; }
    const double d = A[c * n + c];
    ld += log(fabs(d));
    A[c * n + c] = 1.0;
    for (int j = 0; j < n; j++) A[c * n + j] /= d;
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      const double f = A[r * n + c];
      if (f == 0.0) continue;
      A[r * n + c] = 0.0;
      for (int j = 0; j < n; j++) A[r * n + j] -= f * A[c * n + j];
    }
  }
  if (!ok) for (int c = n - 1; c >= 0; c--) if (piv[c] != c) for (int r = 0; r < n; r++) { double tmp = A[r * n + c]; A[r * n + c] = A[r * n + piv[c]]; A[r * n + piv[c]] = tmp; }
  free(piv);
  if (logdet) *logdet = ld;
  return ok;
}

static double fmllr_auxf(const double *W, const double *stats, int D) {
  const int D1 = D + 1, NP = D1 * (D1 + 1) / 2;
  const double beta = stats[0], *K = stats + 1, *G = K + (size_t)D * D1;
  double *A = malloc(sizeof(double) * D * D), ld = 0.0;
  for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) A[i * D + j] = W[i * D1 + j];
  mat_invert(A, D, &ld);
  free(A);
  double obj = beta * ld;
  for (int i = 0; i < D * D1; i++) obj += W[i] * K[i];
  for (int d = 0; d < D; d++) {
    const double *Gd = G + (size_t)d * NP, *w = W + d * D1;
    double q = 0.0;
    for (int i = 0; i < D1; i++) {
      double r = 0.0;
      for (int j = 0; j < D1; j++) r += (j <= i ? Gd[i * (i + 1) / 2 + j] : Gd[j * (j + 1) / 2 + i]) * w[j];
      q += r * w[i];
    }
    obj -= 0.5 * q;
  }
  return obj;
}

/* W: [D][D+1] float, in = starting transform (unit for gmm-est-fmllr), out = estimate.  Returns the objective improvement
 * (0 and W untouched when beta <= min_count or the objective did not increase). */
ORC_API double orc_fmllr_update(const double *stats, int D, int num_iters, double min_count, float *W) {
  const int D1 = D + 1, NP = D1 * (D1 + 1) / 2;
  const double beta = stats[0], *K = stats + 1, *G = K + (size_t)D * D1;
  if (!(beta > min_count)) return 0.0;
  double *invG = malloc(sizeof(double) * D * D1 * D1);
  for (int d = 0; d < D; d++) {
    double *M = invG + (size_t)d * D1 * D1;
    const double *Gd = G + (size_t)d * NP;
    for (int i = 0; i < D1; i++) for (int j = 0; j < D1; j++) M[i * D1 + j] = j <= i ? Gd[i * (i + 1) / 2 + j] : Gd[j * (j + 1) / 2 + i];
    mat_invert(M, D1, NULL);
  }
  double *Wo = malloc(sizeof(double) * D * D1), *Wn = malloc(sizeof(double) * D * D1), *cof = malloc(sizeof(double) * D * D);
  double *c = malloc(sizeof(double) * D1), *cg = malloc(sizeof(double) * D1);
  for (int i = 0; i < D * D1; i++) Wo[i] = Wn[i] = W[i];
  const double old_objf = (float)fmllr_auxf(Wo, stats, D);
  for (int it = 0; it < num_iters; it++) {
    for (int row = 0; row < D; row++) {
      const double *iG = invG + (size_t)row * D1 * D1, *k = K + (size_t)row * D1;
      for (int i = 0; i < D; i++) for (int j = 0; j < D; j++) cof[i * D + j] = Wn[j * D1 + i];   /* A^T */
      mat_invert(cof, D, NULL);
      for (int j = 0; j < D; j++) c[j] = cof[row * D + j];
      c[D] = 0.0;
      for (int i = 0; i < D1; i++) { double r = 0.0; for (int j = 0; j < D1; j++) r += iG[i * D1 + j] * c[j]; cg[i] = r; }
      double e1 = 0.0, e2 = 0.0;
      for (int i = 0; i < D1; i++) { e1 += cg[i] * c[i]; e2 += cg[i] * k[i]; }
      const double discr = sqrt(e2 * e2 + 4 * e1 * beta);
      const double a1 = (-e2 + discr) / (2 * e1), a2 = (-e2 - discr) / (2 * e1);
      const double f1 = beta * log(fabs(a1 * e1 + e2)) - 0.5 * a1 * a1 * e1, f2 = beta * log(fabs(a2 * e1 + e2)) - 0.5 * a2 * a2 * e1;
      const double alpha = f1 > f2 ? a1 : a2;
      for (int i = 0; i < D1; i++) c[i] = alpha * c[i] + k[i];
      for (int i = 0; i < D1; i++) { double r = 0.0; for (int j = 0; j < D1; j++) r += iG[i * D1 + j] * c[j]; Wn[row * D1 + i] = r; }
    }
  }
  const double new_objf = (float)fmllr_auxf(Wn, stats, D), impr = new_objf - old_objf;
  double ret = 0.0;
  const int approx_equal = fabs(new_objf - old_objf) <= 0.001 * (fabs(new_objf) + fabs(old_objf));
  if (!(impr < 0.0 && !approx_equal)) { for (int i = 0; i < D * D1; i++) W[i] = (float)Wn[i]; ret = impr; }
  free(invG); free(Wo); free(Wn); free(cof); free(c); free(cg);
  return ret;
}

/* sizes of the structs, so the ctypes mirror can assert it matches */
ORC_API int orc_sizeof_mfcc_opts(void) { return (int)sizeof(orc_mfcc_opts); }
ORC_API int orc_sizeof_gmm(void) { return (int)sizeof(orc_gmm); }
ORC_API int orc_sizeof_fst(void) { return (int)sizeof(orc_fst); }
