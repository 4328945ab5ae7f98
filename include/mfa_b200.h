/*
 * mfa_b200.h -- C ABI of the B200-native engine behind MFA's alignment hot path.
 *
 * The reference has no C/FFI boundary of its own: the path runs inside kalpy's pybind11 objects
 * (SURVEY.md section 8b).  Each entry point below names the kalpy call it replaces and the MFA
 * call site (paths relative to /root/reference/montreal_forced_aligner) that reaches it.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every function returns MFA_OK (0) or a negative error code; mfa_last_error() gives the text
 *     (thread-local).  No exceptions cross the ABI.
 *   - `where` says whether the caller's DATA buffers are host (MFA_HOST: the call does the H2D/D2H
 *     copies on the engine's stream and returns after the results are in the host buffers) or
 *     device (MFA_DEVICE: pointers are cudaMalloc'd / torch CUDA storage; the call only enqueues
 *     work on the engine's stream -- use mfa_engine_sync()).
 *     Small index arrays (offsets, utt2spk, option structs) are ALWAYS host pointers.
 *   - one engine = one device + one CUDA stream + grow-only workspaces; not shared across threads.
 *   - there is NO CPU fallback: device entry points fail with MFA_ERR_CUDA when no GPU is usable.
 */
#ifndef MFA_B200_H_
#define MFA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFA_API __attribute__((visibility("default")))

#define MFA_OK 0
#define MFA_ERR_INVALID (-1)
#define MFA_ERR_CUDA (-2)
#define MFA_ERR_NOMEM (-3)
#define MFA_ERR_UNSUPPORTED (-4)
#define MFA_ERR_GRAPH (-5)

#define MFA_HOST 0
#define MFA_DEVICE 1

/* per-utterance alignment status (AlignUtteranceWrapper outcomes, decoder/decoder-wrappers.cc) */
#define MFA_ALIGN_OK 0
#define MFA_ALIGN_RETRIED 1
#define MFA_ALIGN_NO_FINAL 2
#define MFA_ALIGN_EMPTY_GRAPH 3
#define MFA_ALIGN_ZERO_FRAMES 4
/* the utterance's training graph has more than 65 534 states or arcs (3-4 minutes of continuous speech): the packed graph views index
 * states and arcs with 16 bits.  The utterance is skipped like any other failure (kalpy returns None), the rest of the batch is aligned;
 * MFA's own segmentation (TextGrid intervals, VAD) keeps utterances far below this. */
#define MFA_ALIGN_GRAPH_TOO_LARGE 5

typedef struct mfa_engine mfa_engine;
typedef struct mfa_model mfa_model;
typedef struct mfa_graph_compiler mfa_graph_compiler;
typedef struct mfa_fst_batch mfa_fst_batch;
typedef struct mfa_graphs mfa_graphs;

MFA_API const char *mfa_last_error(void);
MFA_API int mfa_abi_version(void);

/* ---- engine ------------------------------------------------------------------------------ */
MFA_API int mfa_engine_create(int device, mfa_engine **out);
MFA_API int mfa_engine_destroy(mfa_engine *e);
MFA_API int mfa_engine_sync(mfa_engine *e);
MFA_API void *mfa_engine_stream(mfa_engine *e); /* cudaStream_t, for torch interop */
MFA_API int mfa_engine_sm_count(mfa_engine *e);
/* Experiment / test switches of one engine (integers).  Initial values come from the environment variables MFA_<NAME IN UPPER CASE>,
 * read once inside mfa_engine_create; afterwards only these calls change them.  Names: vit_band, vit_maxgroups, vit_graph_smem,
 * vit_nw2_kb, vit_carveout, vit_carveout_band, vit_prio (creation time only), pipeline_split, acc_impl, tc_k96 (takes effect when a
 * model's operand images are built), tc_poly, mfcc_generic, trace.  Unknown names fail with MFA_ERR_INVALID. */
MFA_API int mfa_engine_set_option(mfa_engine *e, const char *name, int value);
MFA_API int mfa_engine_get_option(mfa_engine *e, const char *name, int *value);
/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
MFA_API int64_t mfa_engine_launch_count(mfa_engine *e);
/* Cumulative number of utterances the band Viterbi kernel handed to the sparse kernel (live window wider than the band). */
MFA_API int64_t mfa_engine_band_fallbacks(mfa_engine *e);
/* CUDA-event timing (on the engine stream) of the K2 launches issued by the most recent API call that ran K2:
 * total milliseconds, number of K2 kernel launches and frame rows they covered (padding included). */
MFA_API int mfa_engine_gmm_timing(mfa_engine *e, float *total_ms, int64_t *n_launches, int64_t *n_rows);
/* useful FLOPs of those launches: 2*(2*dim+1) per (frame, Gaussian) actually scored.  The fused pipeline scores, per utterance,
 * only the pdfs its graph references (what Kaldi's decodable evaluates lazily), so this is less than frames x all Gaussians. */
MFA_API int mfa_engine_gmm_flops(mfa_engine *e, double *useful_flops);
/* FLOPs the tensor-core launches of the last call ISSUED: 3 fp16 products x 2 x K x 128 columns x 128 frames per (frame tile, Gaussian
 * tile), padding columns / rows / K included; 0 when the last call scored with a CUDA-core kernel. */
MFA_API int mfa_engine_gmm_issued_flops(mfa_engine *e, double *issued_flops);
/* CUDA-event time (ms, engine stream) the last mfa_align_pcm call spent per stage:
 * ms4[0] = K1 MFCC + CMVN statistics (host-buffer calls: including the piecewise PCM upload it overlaps with),
 * ms4[1] = feature finalisation, ms4[2] = K2 log-likelihoods, ms4[3] = K3 Viterbi (all size classes, fork to join). */
MFA_API int mfa_engine_stage_timing(mfa_engine *e, float *ms4);

/* ---- K1: MFCC.  Replaces kalpy MfccComputer.compute_mfccs_for_export
 *      (corpus/features.py:235, online/alignment.py:83).  Options = FeatureConfigMixin.mfcc_options
 *      (corpus/features.py:780-820); dither is not supported (parity runs force dither=0). */
typedef struct {
  float sample_frequency, frame_length_ms, frame_shift_ms, preemph_coeff, low_freq, high_freq;
  float cepstral_lifter, energy_floor;
  int32_t num_mel_bins, num_ceps, use_energy, raw_energy, snip_edges, remove_dc_offset;
} mfa_mfcc_opts;

MFA_API int64_t mfa_mfcc_num_frames(const mfa_mfcc_opts *o, int64_t num_samples);
/* pcm: concatenated int16 samples of n_utts utterances; sample_off[n_utts+1], frame_off[n_utts+1]
 * (frame_off[u+1]-frame_off[u] == mfa_mfcc_num_frames(len_u)); out: [frame_off[n_utts]][num_ceps] f32 */
MFA_API int mfa_mfcc(mfa_engine *e, const mfa_mfcc_opts *o, const int16_t *pcm, const int64_t *sample_off,
                     int32_t n_utts, const int64_t *frame_off, float *out, int where);

/* ---- CMVN.  Replaces CmvnComputer.export_cmvn / compute_cmvn_from_features and ApplyCmvn
 *      (corpus/acoustic_corpus.py:1336, command_line/align_one.py:168,183).
 *      stats: [n_spk][2][dim+1] f64 (Kaldi layout, [0][dim] = count). norm_vars=false. */
MFA_API int mfa_cmvn_stats(mfa_engine *e, const float *feats, int32_t dim, const int64_t *frame_off,
                           const int32_t *utt2spk, int32_t n_utts, int32_t n_spk, double *stats, int where);
MFA_API int mfa_cmvn_apply(mfa_engine *e, float *feats, int32_t dim, const int64_t *frame_off,
                           const int32_t *utt2spk, int32_t n_utts, int32_t n_spk, const double *stats, int where);

/* ---- a5: feature finalisation.  Replaces FeatureArchive(deltas | splices+lda, fMLLR)
 *      (db.py:2101-2136; op order at alignment/multiprocessing.py:1287-1304).
 *      mode 0: copy; 1: add deltas (order 2, window 2); 2: splice +-ctx then `lda` [lda_rows][lda_cols]
 *      (lda_cols == in_dim*(2ctx+1) or +1 for an affine column).  Then, if fmllr != NULL, per-speaker
 *      affine [n_spk][D][D+1].  cmvn_stats != NULL fuses the mean subtraction in front.
 *      All small matrices (lda, fmllr, cmvn_stats) are host pointers. */
typedef struct {
  int32_t mode, in_dim, splice_ctx, lda_rows, lda_cols, n_spk;
  const float *lda;
  const float *fmllr;
  const double *cmvn_stats;
} mfa_feat_opts;
MFA_API int32_t mfa_feat_out_dim(const mfa_feat_opts *o);
MFA_API int mfa_features(mfa_engine *e, const mfa_feat_opts *o, const float *in, const int64_t *frame_off,
                         const int32_t *utt2spk, int32_t n_utts, float *out, int where);

/* ---- acoustic model.  Replaces read_gmm_model / AmDiagGmm + TransitionModel tables
 *      (alignment/multiprocessing.py:1393; acoustic_modeling/base.py:299 upstream). */
typedef struct {
  int32_t dim, num_pdfs, num_gauss, num_tids;
  const int32_t *pdf_off;      /* [num_pdfs+1] Gaussians of pdf j = rows pdf_off[j]..pdf_off[j+1]-1 */
  const float *gconsts;        /* [num_gauss] */
  const float *means_invvars;  /* [num_gauss][dim] */
  const float *inv_vars;       /* [num_gauss][dim] */
  const int32_t *tid2pdf;      /* [num_tids+1], index 0 unused */
  const float *weights;        /* [num_gauss] mixture weights, or NULL (only needed to read an updated model back: mfa_model_read) */
} mfa_model_desc;
MFA_API int mfa_model_create(mfa_engine *e, const mfa_model_desc *d, mfa_model **out);
MFA_API int mfa_model_destroy(mfa_model *m);
/* GmmAligner.boost_silence (alignment/multiprocessing.py:803-815): gconst += log(factor) for pdfs */
MFA_API int mfa_model_boost_pdfs(mfa_model *m, float factor, const int32_t *pdfs, int32_t n);

/* ---- N3 / a10: the M-step on the device.  Replaces, after the accumulators of all jobs have been summed (here: the NCCL all-reduce of
 *      the block behind mfa_acc_device_ptr), `tm.mle_update(transition_accs)` and `am.mle_update(gmm_accs, mixup=, power=)` of
 *      AcousticModelTrainingMixin.acc_stats (acoustic_modeling/base.py:319-338 upstream; monophone.py:275-296): Kaldi MleDiagGmmUpdate /
 *      MleAmDiagGmmUpdate, AmDiagGmm::SplitByCount + DiagGmm::Split, TransitionModel::MleUpdate.  The model is updated IN PLACE from its
 *      own accumulator block (which is released: call mfa_acc_zero before the next mfa_acc_stats; mfa_acc_size changes with the number
 *      of Gaussians), and the K2 operand images are rebuilt on the device.  No accumulator leaves the GPU. */
typedef struct {
  int32_t num_tstates;
  const int32_t *tstate_first_tid;  /* [num_tstates+2], 1-based transition-states -> first transition-id */
  const int32_t *self_loop_tid;     /* [num_tstates+1] self-loop transition-id of the state, 0 if none */
  const float *log_probs;           /* [num_tids+1] */
} mfa_trans_desc;
MFA_API int mfa_model_set_transitions(mfa_model *m, const mfa_trans_desc *d);
typedef struct {
  double min_gaussian_occupancy;    /* Kaldi 10 (MFA: 3 at monophone iteration 0, monophone.py:282) */
  double min_gaussian_weight;       /* 1e-5 */
  double min_variance;              /* 1e-3 */
  int32_t remove_low_count_gaussians; /* 1 */
  int32_t mixup;                    /* target total number of Gaussians (SplitByCount), 0 = no mix-up */
  float power;                      /* 0.25 */
  float min_count;                  /* 20 */
  float perturb_factor;             /* 0.01 */
  int32_t update_transitions;       /* 1: TransitionModel::MleUpdate from the transition counts of the block */
  float transition_floor;           /* 0.01 */
  float transition_mincount;        /* 5 */
  uint64_t seed;                    /* of the counter-based normal draws of the mix-up perturbation */
} mfa_mle_opts;
typedef struct {
  double gmm_objf_impr, gmm_count;  /* summed over the pdfs that lost no component (Kaldi reports no change for the others) */
  double trans_objf_impr, trans_count;
  double tot_like, tot_frames;      /* of the accumulation pass that fed this update */
  int64_t variance_floored;
  int32_t num_gauss_before, num_gauss_after, num_removed, num_split;
  int32_t layout_changed;           /* 1: some pdf's component count changed (per-utterance K2 tile plans are rebuilt on next use) */
} mfa_mle_result;
MFA_API int mfa_model_mle_update(mfa_engine *e, mfa_model *m, const mfa_mle_opts *o, mfa_mle_result *res);
/* Training loops: allocate everything the M-step and the accumulators will need for models of up to `max_gauss` Gaussians NOW (both
 * parameter sets the M-step swaps between, the accumulator block), so that no iteration calls cudaMalloc / cudaFree -- on some
 * platforms those calls take tens to hundreds of milliseconds.  max_gauss >= current Gaussians + mix-up target + 1 covers any update. */
MFA_API int mfa_model_reserve(mfa_engine *e, mfa_model *m, int64_t max_gauss);
MFA_API int mfa_model_num_gauss(const mfa_model *m);
/* device -> host copy of the current parameters (any pointer may be NULL): pdf_off[num_pdfs+1], weights / gconsts [num_gauss],
 * means_invvars / inv_vars [num_gauss][dim], log_probs[num_tids+1] -- what write_gmm_model needs ({it}.mdl, acoustic_modeling/base.py). */
MFA_API int mfa_model_read(mfa_engine *e, mfa_model *m, int32_t *pdf_off, float *weights, float *gconsts, float *means_invvars,
                           float *inv_vars, float *log_probs);

/* ---- K2: all-pdf frame log-likelihoods.  Replaces DecodableAmDiagGmmScaled / gmm_compute_likes
 *      (inside GmmAligner; alignment/multiprocessing.py:1415).  out: [n_frames][num_pdfs] f32,
 *      UNSCALED log-likelihoods.  impl: 0 = auto (tcgen05 tensor-core kernel), 1 = fp32 CUDA-core
 *      kernel (exact-order reference kernel used for cross-checks), 2 = tcgen05. */
MFA_API int mfa_gmm_loglikes(mfa_engine *e, mfa_model *m, const float *feats, int64_t n_frames, float *out,
                             int where, int impl);

/* ---- N1: training-graph compiler (host C++).  Replaces TrainingGraphCompiler.compile_fst /
 *      export_graphs (alignment/multiprocessing.py:537-571; online/alignment.py:77-96). */
typedef struct {
  /* topology (hmm/hmm-topology.h), flattened */
  int32_t num_phones;               /* max phone id + 1 */
  const int32_t *phone2entry;       /* [num_phones] topology entry or -1 */
  int32_t num_entries;
  const int32_t *entry_state_off;   /* [num_entries+1] into the hmm-state arrays */
  const int32_t *state_fwd_class;   /* [n_hmm_states] forward pdf-class (-1 = non-emitting) */
  const int32_t *state_self_class;  /* [n_hmm_states] self-loop pdf-class */
  const int32_t *state_trans_off;   /* [n_hmm_states+1] */
  const int32_t *trans_dst;         /* [n_trans] destination hmm state within the entry */
  /* transition model tuples (hmm/transition-model.h) */
  int32_t num_tstates;
  const int32_t *tuples;            /* [num_tstates][4] phone, hmm_state, fwd_pdf, self_loop_pdf */
  const int32_t *tstate_first_tid;  /* [num_tstates+2], 1-based transition-state ids */
  /* context dependency (tree/context-dep.h) */
  int32_t ctx_width, central_pos;   /* N, P : (1,0) or (3,1) */
  int32_t num_tree_nodes, tree_root;
  const int32_t *tree_nodes;        /* [num_tree_nodes][4] type(0 CE,1 SE,2 TE), key, a, b */
  const int32_t *tree_aux_off;      /* [num_tree_nodes+1] */
  const int32_t *tree_aux;          /* SE: sorted yes-set; TE: children (-1 = NULL) */
} mfa_hmm_desc;

typedef struct {
  int32_t num_words;
  const int32_t *word_pron_off;   /* [num_words+1] pronunciations of word id w */
  const int32_t *pron_phone_off;  /* [num_prons+1] */
  const int32_t *pron_phones;
  const float *pron_cost;         /* -log pronunciation probability */
  const float *pron_sil_after_cost;     /* per pron, or NULL -> sil_cost */
  const float *pron_nonsil_after_cost;  /* per pron, or NULL -> nonsil_cost */
  const float *pron_sil_before_cost;    /* per pron correction, or NULL -> 0 */
  const float *pron_nonsil_before_cost; /* per pron correction, or NULL -> 0 */
  int32_t sil_phone;
  float sil_cost, nonsil_cost;            /* -log p_sil, -log(1-p_sil) */
  float init_sil_cost, init_nonsil_cost;  /* -log p_init_sil, -log(1-p_init_sil) */
  float final_sil_cost, final_nonsil_cost;
} mfa_lexicon_desc;

MFA_API int mfa_graph_compiler_create(const mfa_hmm_desc *h, const mfa_lexicon_desc *l, mfa_graph_compiler **out);
MFA_API int mfa_graph_compiler_destroy(mfa_graph_compiler *c);
/* words: concatenated word ids of n_utts transcripts, word_off[n_utts+1]. */
MFA_API int mfa_graph_compile(mfa_graph_compiler *c, const int32_t *words, const int64_t *word_off, int32_t n_utts,
                              int32_t n_threads, mfa_fst_batch **out);

/* arc-list FST batches (also the import path for FstArchive: OpenFst VectorFst<StdArc> arrays) */
MFA_API int mfa_fst_batch_create(int32_t n_utts, const int64_t *state_off, const int64_t *arc_off, const int32_t *start,
                                 const float *finals, const int32_t *src, const int32_t *dst, const int32_t *ilabel,
                                 const int32_t *olabel, const float *weight, mfa_fst_batch **out);
MFA_API int mfa_fst_batch_destroy(mfa_fst_batch *b);
/* The state / arc body of a binary OpenFst VectorFst<StdArc> (what kalpy's FstArchive holds, alignment/multiprocessing.py:831): per state a
 * float final weight and an int64 arc count, then 16-byte arcs {ilabel, olabel, weight, nextstate}.  `body` points behind the FST header
 * (the host parses that: magic, type strings, flags, start, number of states).  scan: number of arcs and bytes of the body (error when
 * it does not fit into `len`); fill: the arrays of mfa_fst_batch_create's arc-list form (finals[n_states], the rest [n_arcs]). */
MFA_API int mfa_fst_body_scan(const uint8_t *body, int64_t len, int64_t n_states, int64_t *n_arcs, int64_t *n_bytes);
MFA_API int mfa_fst_body_fill(const uint8_t *body, int64_t n_states, float *finals, int32_t *src, int32_t *dst, int32_t *ilabel,
                              int32_t *olabel, float *weight);
MFA_API int mfa_fst_batch_sizes(const mfa_fst_batch *b, int32_t *n_utts, int64_t *n_states, int64_t *n_arcs);
MFA_API int mfa_fst_batch_export(const mfa_fst_batch *b, int64_t *state_off, int64_t *arc_off, int32_t *start,
                                 float *finals, int32_t *src, int32_t *dst, int32_t *ilabel, int32_t *olabel,
                                 float *weight);

/* ---- a11: equal alignment (host).  Replaces kalpy gmm_align_equal -> Kaldi EqualAlign + GetLinearSymbolSequence, called by
 *      MonoAlignEqualFunction._run (acoustic_modeling/monophone.py:108) for iteration 0 of monophone training.  One random
 *      self-loop-free path per graph (seeds[u]: Kaldi's align-equal-compiled seeds srand() with StringHasher(utterance id);
 *      the draws restate glibc rand()), the remaining frames spread evenly over the path's self-loops.  Outputs as mfa_align
 *      (ali at frame_off, words at word_off with capacity word_off[u+1]-word_off[u], status MFA_ALIGN_OK / NO_FINAL = could not
 *      match the length / EMPTY_GRAPH / ZERO_FRAMES).  num_retries <= 0 -> 10 (Kaldi's default). */
MFA_API int mfa_equal_align(const mfa_fst_batch *b, const int64_t *frame_off, const uint32_t *seeds, int32_t num_retries,
                            int32_t *ali, int32_t *words, const int64_t *word_off, int32_t *num_words, int32_t *status);
/* the first n values of rand() after srand(seed) as restated for mfa_equal_align (tests compare with libc) */
MFA_API int mfa_rand_sequence(uint32_t seed, int32_t n, int32_t *out);

/* decoder-ready packing: AddTransitionProbs (tid_cost[tid] = -scaled log prob, hmm/hmm-utils.cc) folded into the
 * arc weights, arcs grouped by destination, per-utterance local pdf lists.  Host-resident; uploaded lazily. */
MFA_API int mfa_graphs_pack(const mfa_fst_batch *b, const float *tid_cost, const int32_t *tid2pdf, int32_t num_tids,
                            mfa_graphs **out);
MFA_API int mfa_graphs_destroy(mfa_graphs *g);
/* AddTransitionProbs again, on the device, from the model's CURRENT transition log-probabilities (after mfa_model_mle_update): what
 * GmmAligner does per graph at align time when it reads the re-estimated model (alignment/multiprocessing.py:814-853). */
MFA_API int mfa_graphs_set_transitions(mfa_engine *e, mfa_graphs *g, mfa_model *m, float transition_scale, float self_loop_scale);
MFA_API int mfa_graphs_max_words(const mfa_graphs *g, int32_t *max_words /* [n_utts] upper bound on olabels per path */);
/* per-utterance prefix offsets ([n_utts+1] each; any pointer may be NULL): states, arcs, distinct pdfs referenced */
MFA_API int mfa_graphs_offsets(const mfa_graphs *g, int64_t *state_off, int64_t *arc_off, int64_t *pdf_off);
/* Band view of the packed graphs (the layout the primary Viterbi kernel runs on; see csrc/viterbi_band.cu): per utterance
 * band_ok[n] (0: sparse kernel only), start[n] and maxback[n] in band numbering; per state (offsets = state_off) the packed
 * word first-in-arc | in-degree << 16 | forward reach << 24 and the original state id; per arc (offsets = arc_off), grouped by
 * DESTINATION band state: source band state | local pdf << 16, and the index of the same arc in by-source order.  Any pointer
 * may be NULL.  Host-only (tests / inspection). */
MFA_API int mfa_graphs_band_view(const mfa_graphs *g, int32_t *band_ok, int32_t *start, int32_t *maxback, uint32_t *state_word,
                                 uint16_t *orig_state, uint32_t *arc_word, uint16_t *arc_index);

/* ---- K3: batched beam Viterbi.  Replaces GmmAligner.align_utterance / export_alignments ->
 *      AlignUtteranceWrapper + FasterDecoder (alignment/multiprocessing.py:846-853;
 *      online/alignment.py:97-107).  loglikes: [frame_off[n]][num_pdfs] from mfa_gmm_loglikes.
 *      Outputs: ali[frame_off[n]] transition-ids; per_frame[frame_off[n]] unscaled log-likes of the
 *      aligned pdf; words[word_off[u]..] olabels (word_off from the caller, capacity per utt),
 *      num_words[n]; total_like[n] = -(path cost)/acoustic_scale; status[n] (MFA_ALIGN_*). */
typedef struct {
  float acoustic_scale, beam, retry_beam, beam_delta;
  int32_t min_active;
} mfa_align_opts;
MFA_API int mfa_align(mfa_engine *e, mfa_model *m, mfa_graphs *g, const mfa_align_opts *o, const float *loglikes,
                      const int64_t *frame_off, int32_t n_utts, int32_t *ali, float *per_frame, int32_t *words,
                      const int64_t *word_off, int32_t *num_words, float *total_like, int32_t *status, int where);

/* ---- features -> alignments: K2 (per utterance only the pdfs of its graph) + K3 on FINAL features [frame_off[n]][dim], no log-likelihood
 *      matrix crossing the boundary.  This is what GmmAligner.align_utterance / export_alignments run per batch
 *      (alignment/multiprocessing.py:846-853): kalpy's decodable evaluates the GMMs lazily inside the decoder, so a drop-in must not
 *      materialise frames x all pdfs either.  gmm_impl as in mfa_pipeline_opts; workspace_bytes 0 = 8 GiB.  Outputs as mfa_align. */
MFA_API int mfa_align_feats(mfa_engine *e, mfa_model *m, mfa_graphs *g, const mfa_align_opts *o, const float *feats,
                            const int64_t *frame_off, int32_t n_utts, int32_t gmm_impl, int64_t workspace_bytes, int32_t *ali,
                            float *per_frame, int32_t *words, const int64_t *word_off, int32_t *num_words, float *total_like,
                            int32_t *status, int where);

/* ---- fused hot path: PCM -> MFCC -> CMVN -> features -> loglikes -> Viterbi, chunked so that the
 *      log-likelihood and back-pointer buffers stay inside `workspace_bytes` of HBM.
 *      This is AlignFunction._run's per-job loop (alignment/multiprocessing.py:791-863) on one GPU. */
typedef struct {
  mfa_mfcc_opts mfcc;
  mfa_feat_opts feat;        /* cmvn_stats may be NULL: then per-speaker stats are computed on device */
  mfa_align_opts align;
  int32_t apply_cmvn;        /* 1: per-speaker CMVN (stats computed on device unless feat.cmvn_stats given) */
  int32_t gmm_impl;          /* 0 = auto: tcgen05, per utterance only the pdfs of its graph; 1 = fp32 CUDA-core kernel, all pdfs;
                                2 = tcgen05, all pdfs (dense) */
  int64_t workspace_bytes;   /* 0 = default (8 GiB) */
} mfa_pipeline_opts;
MFA_API int mfa_align_pcm(mfa_engine *e, mfa_model *m, mfa_graphs *g, const mfa_pipeline_opts *o, const int16_t *pcm,
                          const int64_t *sample_off, const int32_t *utt2spk, int32_t n_utts, int32_t n_spk,
                          const int64_t *frame_off, int32_t *ali, float *per_frame, int32_t *words,
                          const int64_t *word_off, int32_t *num_words, float *total_like, int32_t *status, int where);

/* ---- K4: GMM accumulator statistics.  Replaces GmmStatsAccumulator.accumulate_stats /
 *      AccumAmDiagGmm.acc_stats (alignment/multiprocessing.py:652-666; acoustic_modeling/monophone.py:114-120).
 *      accs layout (f64, device-resident inside the engine until read back):
 *        occ[G] | mean_acc[G][D] | var_acc[G][D] | trans_acc[num_tids+1] | tot_like | tot_frames
 *      mfa_acc_size returns the number of doubles.  frames with ali <= 0 are skipped. */
MFA_API int64_t mfa_acc_size(const mfa_model *m);
MFA_API int mfa_acc_zero(mfa_engine *e, mfa_model *m);
MFA_API int mfa_acc_stats(mfa_engine *e, mfa_model *m, const float *feats, const int32_t *ali, int64_t n_frames,
                          int where);
/* device pointer to the accumulator block (for an NCCL all-reduce by the host layer) */
MFA_API double *mfa_acc_device_ptr(mfa_engine *e, mfa_model *m);
MFA_API int mfa_acc_read(mfa_engine *e, mfa_model *m, double *host_out);
/* host block (same layout) -> device accumulators: for callers that summed kalpy-style accumulator objects on the host */
MFA_API int mfa_acc_write(mfa_engine *e, mfa_model *m, const double *host_in);

/* ---- K5 (N2): per-speaker fMLLR statistics.  Replaces the accumulation inside kalpy FmllrComputer.export_transforms
 *      (CalcFmllrFunction._run, corpus/features.py:460-548; Kaldi gmm-est-fmllr / gmm-est-fmllr-gpost + weight-silence-post).
 *      post_model: the model the component posteriors come from (the alignment model in the two-model case; NULL = m);
 *      m: the model whose means / variances enter the statistics; both are evaluated on the same `feats`.
 *      tid_weight: host [num_tids+1] frame weight per transition-id (silence_weight for silence phones, else 1; NULL = 1).
 *      stats: [n_spk][mfa_fmllr_stats_size(dim)] f64 = beta | K[D][D+1] | G[D][(D+1)(D+2)/2] (G_d: packed lower triangle,
 *      row-major, Kaldi SpMatrix order).  dim <= 63 (the kernel stages [x | 1] in 64-wide rows).  Transform update: mfa_fmllr_update. */
MFA_API int64_t mfa_fmllr_stats_size(int32_t dim);
MFA_API int mfa_fmllr_acc(mfa_engine *e, mfa_model *post_model, mfa_model *m, const float *feats, const int32_t *ali,
                          const float *tid_weight, const int64_t *frame_off, const int32_t *utt2spk, int32_t n_utts,
                          int32_t n_spk, double *stats, int where);
/* fMLLR transform update from those statistics, one CTA per speaker: Kaldi ComputeFmllrMatrixDiagGmmFull started from the unit
 * transform (gmm-est-fmllr): num_iters (<= 0 -> 40) sweeps of the row update, accepted only if the auxiliary function did not
 * decrease; speakers with beta <= min_count (Kaldi: 500) keep the unit transform.  stats / transforms [n_spk][dim][dim+1] f32 follow
 * `where`; objf_impr[n_spk] and count[n_spk] (either may be NULL) are host pointers. */
MFA_API int mfa_fmllr_update(mfa_engine *e, const double *stats, int32_t dim, int32_t n_spk, int32_t num_iters, double min_count,
                             float *transforms, double *objf_impr, double *count, int where);

#ifdef __cplusplus
}
#endif
#endif /* MFA_B200_H_ */
