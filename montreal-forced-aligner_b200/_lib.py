"""ctypes binding of libmfa_b200.so (the C ABI declared in include/mfa_b200.h).

The product path fails loudly when the library is missing: there is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# read by the CUDA runtime when the context is created: an engine drives ~17 streams and MFA runs several jobs per GPU, the default of 8
# hardware queues makes unrelated streams wait for one another (csrc/engine.cu sets the same default when the library is loaded)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
LIB_PATH = os.environ.get("MFA_B200_LIB") or os.path.join(HERE, "libmfa_b200.so")   # MFA_B200_LIB: a development build (tools/k2_experiment.py)

MFA_HOST, MFA_DEVICE = 0, 1
ALIGN_STATUS = {0: "OK", 1: "RETRIED", 2: "NO_FINAL", 3: "EMPTY_GRAPH", 4: "ZERO_FRAMES", 5: "GRAPH_TOO_LARGE"}


class MfaError(RuntimeError):
    pass


class MfccOpts(C.Structure):
    _fields_ = [
        ("sample_frequency", C.c_float), ("frame_length_ms", C.c_float), ("frame_shift_ms", C.c_float),
        ("preemph_coeff", C.c_float), ("low_freq", C.c_float), ("high_freq", C.c_float),
        ("cepstral_lifter", C.c_float), ("energy_floor", C.c_float),
        ("num_mel_bins", C.c_int32), ("num_ceps", C.c_int32), ("use_energy", C.c_int32),
        ("raw_energy", C.c_int32), ("snip_edges", C.c_int32), ("remove_dc_offset", C.c_int32),
    ]


class FeatOpts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("in_dim", C.c_int32), ("splice_ctx", C.c_int32), ("lda_rows", C.c_int32),
                ("lda_cols", C.c_int32), ("n_spk", C.c_int32), ("lda", C.c_void_p), ("fmllr", C.c_void_p),
                ("cmvn_stats", C.c_void_p)]


class ModelDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("num_pdfs", C.c_int32), ("num_gauss", C.c_int32), ("num_tids", C.c_int32),
                ("pdf_off", C.c_void_p), ("gconsts", C.c_void_p), ("means_invvars", C.c_void_p),
                ("inv_vars", C.c_void_p), ("tid2pdf", C.c_void_p), ("weights", C.c_void_p)]


class TransDesc(C.Structure):
    _fields_ = [("num_tstates", C.c_int32), ("tstate_first_tid", C.c_void_p), ("self_loop_tid", C.c_void_p), ("log_probs", C.c_void_p)]


class MleOpts(C.Structure):
    _fields_ = [("min_gaussian_occupancy", C.c_double), ("min_gaussian_weight", C.c_double), ("min_variance", C.c_double),
                ("remove_low_count_gaussians", C.c_int32), ("mixup", C.c_int32), ("power", C.c_float), ("min_count", C.c_float),
                ("perturb_factor", C.c_float), ("update_transitions", C.c_int32), ("transition_floor", C.c_float),
                ("transition_mincount", C.c_float), ("seed", C.c_uint64)]


class MleResult(C.Structure):
    _fields_ = [("gmm_objf_impr", C.c_double), ("gmm_count", C.c_double), ("trans_objf_impr", C.c_double), ("trans_count", C.c_double),
                ("tot_like", C.c_double), ("tot_frames", C.c_double), ("variance_floored", C.c_int64), ("num_gauss_before", C.c_int32),
                ("num_gauss_after", C.c_int32), ("num_removed", C.c_int32), ("num_split", C.c_int32), ("layout_changed", C.c_int32)]


class HmmDesc(C.Structure):
    _fields_ = [("num_phones", C.c_int32), ("phone2entry", C.c_void_p), ("num_entries", C.c_int32),
                ("entry_state_off", C.c_void_p), ("state_fwd_class", C.c_void_p), ("state_self_class", C.c_void_p),
                ("state_trans_off", C.c_void_p), ("trans_dst", C.c_void_p), ("num_tstates", C.c_int32),
                ("tuples", C.c_void_p), ("tstate_first_tid", C.c_void_p), ("ctx_width", C.c_int32),
                ("central_pos", C.c_int32), ("num_tree_nodes", C.c_int32), ("tree_root", C.c_int32),
                ("tree_nodes", C.c_void_p), ("tree_aux_off", C.c_void_p), ("tree_aux", C.c_void_p)]


class LexiconDesc(C.Structure):
    _fields_ = [("num_words", C.c_int32), ("word_pron_off", C.c_void_p), ("pron_phone_off", C.c_void_p),
                ("pron_phones", C.c_void_p), ("pron_cost", C.c_void_p), ("pron_sil_after_cost", C.c_void_p),
                ("pron_nonsil_after_cost", C.c_void_p), ("pron_sil_before_cost", C.c_void_p),
                ("pron_nonsil_before_cost", C.c_void_p), ("sil_phone", C.c_int32), ("sil_cost", C.c_float),
                ("nonsil_cost", C.c_float), ("init_sil_cost", C.c_float), ("init_nonsil_cost", C.c_float),
                ("final_sil_cost", C.c_float), ("final_nonsil_cost", C.c_float)]


class AlignOpts(C.Structure):
    _fields_ = [("acoustic_scale", C.c_float), ("beam", C.c_float), ("retry_beam", C.c_float), ("beam_delta", C.c_float),
                ("min_active", C.c_int32)]


class PipelineOpts(C.Structure):
    _fields_ = [("mfcc", MfccOpts), ("feat", FeatOpts), ("align", AlignOpts), ("apply_cmvn", C.c_int32),
                ("gmm_impl", C.c_int32), ("workspace_bytes", C.c_int64)]


# every symbol include/mfa_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "mfa_last_error", "mfa_abi_version", "mfa_engine_create", "mfa_engine_destroy", "mfa_engine_sync", "mfa_engine_stream",
    "mfa_engine_sm_count", "mfa_engine_set_option", "mfa_engine_get_option", "mfa_engine_launch_count", "mfa_engine_band_fallbacks", "mfa_engine_gmm_timing", "mfa_engine_gmm_flops", "mfa_engine_gmm_issued_flops", "mfa_engine_stage_timing", "mfa_mfcc_num_frames", "mfa_mfcc",
    "mfa_cmvn_stats", "mfa_cmvn_apply", "mfa_feat_out_dim", "mfa_features", "mfa_model_create", "mfa_model_destroy",
    "mfa_model_boost_pdfs", "mfa_model_set_transitions", "mfa_model_mle_update", "mfa_model_reserve", "mfa_model_num_gauss", "mfa_model_read",
    "mfa_graphs_set_transitions", "mfa_gmm_loglikes", "mfa_graph_compiler_create", "mfa_graph_compiler_destroy", "mfa_graph_compile",
    "mfa_fst_batch_create", "mfa_fst_batch_destroy", "mfa_fst_body_scan", "mfa_fst_body_fill", "mfa_fst_batch_sizes", "mfa_fst_batch_export", "mfa_graphs_pack",
    "mfa_graphs_destroy", "mfa_graphs_max_words", "mfa_graphs_offsets", "mfa_graphs_band_view", "mfa_align", "mfa_align_feats", "mfa_align_pcm", "mfa_acc_size", "mfa_acc_zero", "mfa_acc_stats",
    "mfa_acc_device_ptr", "mfa_acc_read", "mfa_acc_write", "mfa_equal_align", "mfa_rand_sequence", "mfa_fmllr_stats_size", "mfa_fmllr_acc", "mfa_fmllr_update",
]

_lib = None


def lib():
    """Load the shared library (building it is __graft_entry__.build()'s job, not an import side effect)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MfaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.mfa_last_error.restype = C.c_char_p
        _lib.mfa_mfcc_num_frames.restype = C.c_int64
        _lib.mfa_mfcc_num_frames.argtypes = [C.c_void_p, C.c_int64]
        _lib.mfa_engine_stream.restype = C.c_void_p
        _lib.mfa_engine_launch_count.restype = C.c_int64
        _lib.mfa_engine_band_fallbacks.restype = C.c_int64
        _lib.mfa_acc_size.restype = C.c_int64
        _lib.mfa_fmllr_stats_size.restype = C.c_int64
        _lib.mfa_acc_device_ptr.restype = C.c_void_p
        for name in ("mfa_engine_destroy", "mfa_engine_sync", "mfa_engine_stream", "mfa_engine_sm_count",
                     "mfa_engine_launch_count", "mfa_engine_band_fallbacks", "mfa_model_destroy", "mfa_graph_compiler_destroy", "mfa_fst_batch_destroy",
                     "mfa_graphs_destroy", "mfa_acc_size"):
            getattr(_lib, name).argtypes = [C.c_void_p]
    return _lib


def check(rc: int):
    if rc != 0:
        raise MfaError(f"mfa_b200 error {rc}: {lib().mfa_last_error().decode('utf8', 'replace')}")
