"""MFA's per-job functions on the hot path, re-hosted on the B200 engine.

Mirrors (same names, argument meaning, files written, callback protocol, error behaviour) of the reference's
  MfccFunction                 montreal_forced_aligner/corpus/features.py:162-251      (row a1)
  FinalFeatureFunction         montreal_forced_aligner/corpus/features.py:254-376      (row a4)
  CompileTrainGraphsFunction   montreal_forced_aligner/alignment/multiprocessing.py:386-574 (row a6)
  AlignFunction                montreal_forced_aligner/alignment/multiprocessing.py:668-863 (row a7)
  AccStatsFunction             montreal_forced_aligner/alignment/multiprocessing.py:576-666 (row a9)
  MonoAlignEqualFunction       montreal_forced_aligner/acoustic_modeling/monophone.py:40-139 (row a11)
  CalcFmllrFunction            montreal_forced_aligner/corpus/features.py:423-548          (row N2)
  AlignmentExtractionFunction  montreal_forced_aligner/alignment/multiprocessing.py:1549-1862 (row N4; + export_textgrids)
and of the drivers calc_cmvn (corpus/acoustic_corpus.py:1315-1367, row a3), AlignMixin.align_utterances
(alignment/mixins.py:282-380, row a8) and AcousticModelTrainingMixin.acc_stats (acoustic_modeling/base.py:277-338 upstream,
row a10).  The reference pulls utterances from its database; here a ``Job`` carries them explicitly (the DB is out of scope).
File names follow Job.construct_path (db.py:2212-2236).
"""
from __future__ import annotations

import os
import shutil
import traceback
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import kaldi_io as K
from . import kalpy_compat as KC
from .gmm_update import AccumAmDiagGmm, mle_update
from .lexicon import Lexicon

MetaDict = dict


class MultiprocessingError(Exception):
    """abc.py:939-952: a job function's exception, tagged with the job."""

    def __init__(self, job_name, error_text):
        super().__init__(f"Job {job_name} encountered an error:\n{error_text}")
        self.job_name, self.error_text = job_name, error_text


class NoAlignmentsError(Exception):
    """exceptions.py:493-512: raised when no utterance could be aligned."""

    def __init__(self, num_utterances, beam, retry_beam):
        super().__init__(f"There were no successful alignments for {num_utterances} utterances with beam {beam} / retry beam {retry_beam}; "
                         f"try beam {beam * 10} / retry beam {retry_beam * 10}")


@dataclass
class Utterance:
    id: int
    speaker_id: int
    path: str
    normalized_text: str = ""
    begin: Optional[float] = None
    end: Optional[float] = None
    channel: int = 0
    duration: float = 1.0
    dictionary_id: int = 1
    ignored: bool = False
    alignment_log_likelihood: Optional[float] = None
    num_frames: Optional[int] = None

    @property
    def kaldi_id(self) -> str:  # db_polars.py:2186-2192
        return f"{self.speaker_id}-{self.id}"


@dataclass
class Job:
    """One unit of parallelism = one set of speakers (corpus/base.py:994-1015); maps to one GPU rank in the B200 build."""
    id: int
    utterances: List[Utterance]
    split_directory: Path
    dictionary_ids: List[int] = field(default_factory=lambda: [1])

    def construct_path(self, directory, identifier: str, extension: str, dictionary_id: Optional[int] = None) -> Path:
        if dictionary_id is None:
            return Path(directory) / f"{identifier}.{self.id}.{extension}"
        return Path(directory) / f"{identifier}.{dictionary_id}.{self.id}.{extension}"

    def utts(self, dictionary_id: Optional[int] = None) -> List[Utterance]:
        us = [u for u in self.utterances if not u.ignored and (dictionary_id is None or u.dictionary_id == dictionary_id)]
        return sorted(us, key=lambda u: u.kaldi_id)

    def write_maps(self, dictionary_id: int):
        """utt2spk / spk2utt / per-dictionary feats scp (corpus/multiprocessing.py:403-566)."""
        us = self.utts(dictionary_id)
        with open(self.construct_path(self.split_directory, "utt2spk", "scp", dictionary_id), "w") as f:
            for u in us:
                f.write(f"{u.kaldi_id} {u.speaker_id}\n")
        feats = {k: (p, o) for k, p, o in K.read_scp(self.construct_path(self.split_directory, "feats", "scp"))}
        with open(self.construct_path(self.split_directory, "feats", "scp", dictionary_id), "w") as f:
            for u in us:
                if u.kaldi_id in feats:
                    p, o = feats[u.kaldi_id]
                    f.write(f"{u.kaldi_id} {p}:{o}\n")

    def construct_feature_archive(self, working_directory, dictionary_id: Optional[int] = None, **kwargs) -> KC.FeatureArchive:
        """db.py:2101-2136: lda.mat in the working directory -> splice+LDA, else deltas; trans.*.scp -> fMLLR."""
        working_directory = Path(working_directory)
        fmllr = self.construct_path(working_directory, "trans", "scp", dictionary_id)
        if not fmllr.exists():
            fmllr = self.construct_path(self.split_directory, "trans", "scp", dictionary_id)
        lda = working_directory / "lda.mat"
        feat = self.construct_path(self.split_directory, "feats", "scp", dictionary_id)
        if not feat.exists():
            feat = self.construct_path(self.split_directory, "feats", "scp")
        u2s = self.construct_path(self.split_directory, "utt2spk", "scp", dictionary_id)
        return KC.FeatureArchive(feat, utt2spk_file_name=u2s if u2s.exists() else None,
                                 lda_mat_file_name=lda if lda.exists() else None,
                                 transform_file_name=fmllr if fmllr.exists() else None,
                                 deltas=kwargs.get("uses_deltas", not lda.exists()), splices=lda.exists(),
                                 splice_frames=kwargs.get("splice_context", 3))


def assign_jobs(utterances: Sequence[Utterance], num_jobs: int, split_directory) -> List[Job]:
    """Speakers sorted by utterance count, each to the currently lightest job (corpus/base.py:994-1015).  A speaker never
    spans jobs, so per-speaker CMVN / fMLLR stay job-local -- and GPU-local when jobs map to ranks."""
    by_spk: Dict[int, List[Utterance]] = {}
    for u in utterances:
        by_spk.setdefault(u.speaker_id, []).append(u)
    jobs = [Job(i + 1, [], Path(split_directory)) for i in range(num_jobs)]
    load = [0] * num_jobs
    for spk, us in sorted(by_spk.items(), key=lambda kv: (-len(kv[1]), kv[0])):
        j = int(np.argmin(load))
        jobs[j].utterances.extend(us)
        load[j] += len(us)
    return jobs


# ------------------------------------------------------------------------------------------------ function base
@dataclass
class MfaArguments:
    job_name: int
    job: Job            # stands in for the reference's `session` (data.py:268-284)
    log_path: Optional[Path]


class KaldiFunction:
    """abc.py:915-969: run() wraps _run(); results/progress only through self.callback (ints = progress, tuples = payloads)."""

    def __init__(self, args: MfaArguments):
        self.args = args
        self.job_name = args.job_name
        self.job = args.job
        self.log_path = args.log_path
        self.callback: Callable = lambda x: None

    def run(self, callback: Optional[Callable] = None):
        if callback is not None:
            self.callback = callback
        try:
            self._run()
        except Exception:
            raise MultiprocessingError(self.job_name, traceback.format_exc())

    def _run(self):
        raise NotImplementedError


def run_kaldi_function(function, arguments: Sequence[MfaArguments], progress: Optional[Callable] = None, num_threads: int = 1):
    """utils.py:1505-1642.  num_threads = 1: the jobs of this GPU rank run back to back; > 1: MFA's USE_THREADING mode -- jobs run as
    threads of this process (utils.py:1560-1580), each on its own engine (kalpy_compat.get_engine is per thread), so that one job's
    file / Python work overlaps another's GPU work.  Results are yielded per job in job order; the first job error is re-raised."""
    def one(a):
        results = []
        function(a).run(results.append)
        return results

    if num_threads <= 1 or len(arguments) <= 1:
        per_job = (one(a) for a in arguments)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(num_threads, len(arguments))) as ex:
            per_job = list(ex.map(one, arguments))
    for results in per_job:
        for r in results:
            if isinstance(r, int):
                if progress:
                    progress(r)
            else:
                yield r


# ------------------------------------------------------------------------------------------------ a1: MFCC
@dataclass
class MfccArguments(MfaArguments):
    data_directory: Path
    mfcc_computer: KC.MfccComputer
    pitch_computer: Optional[object] = None


class MfccFunction(KaldiFunction):
    def __init__(self, args: MfccArguments):
        super().__init__(args)
        self.data_directory, self.mfcc_computer = Path(args.data_directory), args.mfcc_computer
        if args.pitch_computer is not None:
            raise KC.MfaError("pitch features are outside the hot path (SURVEY.md section 2a)")

    def _run(self):
        ark = self.job.construct_path(self.data_directory, "feats", "ark")
        scp = self.job.construct_path(self.data_directory, "feats", "scp")
        if ark.exists():   # features.py:202-203: resume
            return
        us = [u for u in self.job.utts() if u.duration >= 0.1]   # features.py:206,223
        B = 256
        with K.ArkWriter(ark, scp) as w:
            for i in range(0, len(us), B):
                chunk = us[i:i + B]
                pcm = [KC.Segment(u.path, u.begin, u.end, u.channel).load_audio() for u in chunk]
                mats = self.mfcc_computer.compute_mfccs_batch(pcm)
                for u, m in zip(chunk, mats):
                    if m.shape[0] == 0:
                        u.ignored = True   # acoustic_corpus.py:880-912: feature failure -> utterance ignored
                        continue
                    w.write_matrix(u.kaldi_id, m, compress=True)
                    self.callback(1)


# ------------------------------------------------------------------------------------------------ a3: CMVN
def calc_cmvn(jobs: Sequence[Job], split_directory) -> Path:
    """acoustic_corpus.py:1315-1367: per-speaker statistics over all raw features -> cmvn.ark/scp, then per-job cmvn.J.scp."""
    split_directory = Path(split_directory)
    all_scp = split_directory / "feats.scp"
    spk2utt: Dict[str, List[str]] = {}
    with open(all_scp, "w") as f:
        for j in jobs:
            for k, p, o in K.read_scp(j.construct_path(split_directory, "feats", "scp")):
                f.write(f"{k} {p}:{o}\n")
            for u in j.utts():
                spk2utt.setdefault(str(u.speaker_id), []).append(u.kaldi_id)
    fa = KC.FeatureArchive(all_scp)
    have = set(fa.keys)
    spk2utt = {s: [k for k in ks if k in have] for s, ks in spk2utt.items()}
    spk2utt = {s: ks for s, ks in spk2utt.items() if ks}
    KC.CmvnComputer().export_cmvn(split_directory / "cmvn.ark", fa, spk2utt, write_scp=True)
    entries = {k: (p, o) for k, p, o in K.read_scp(split_directory / "cmvn.scp")}
    for j in jobs:
        with open(j.construct_path(split_directory, "cmvn", "scp"), "w") as f:
            for s in sorted({str(u.speaker_id) for u in j.utts()}):
                if s in entries:
                    f.write(f"{s} {entries[s][0]}:{entries[s][1]}\n")
    return split_directory / "cmvn.scp"


# ------------------------------------------------------------------------------------------------ a4: final features
@dataclass
class FinalFeatureArguments(MfaArguments):
    data_directory: Path
    uses_cmvn: bool = True
    sliding_cmvn: bool = False
    voiced_only: bool = False
    subsample_feats: int = 0


class FinalFeatureFunction(KaldiFunction):
    def __init__(self, args: FinalFeatureArguments):
        super().__init__(args)
        self.data_directory, self.uses_cmvn = Path(args.data_directory), args.uses_cmvn
        if args.sliding_cmvn or args.voiced_only or args.subsample_feats:
            raise KC.MfaError("sliding CMVN / VAD / subsampling are outside the hot path")

    def _run(self):
        d = self.data_directory
        feats_scp = self.job.construct_path(d, "feats", "scp")
        u2s = d / f"utt2spk.{self.job.id}.scp"
        with open(u2s, "w") as f:
            for u in self.job.utts():
                f.write(f"{u.kaldi_id} {u.speaker_id}\n")
        cmvn = self.job.construct_path(d, "cmvn", "scp")
        fa = KC.FeatureArchive(feats_scp, utt2spk_file_name=u2s, cmvn_file_name=cmvn if self.uses_cmvn else None)
        out_ark = self.job.construct_path(d, "final_features", "ark")
        out_scp = self.job.construct_path(d, "final_features", "scp")
        with K.ArkWriter(out_ark, out_scp) as w:
            for k, m in fa:
                w.write_matrix(k, m, compress=True)   # features.py:356-365: CMVN'd features are re-compressed
                self.callback(1)
        fa.close()
        os.replace(out_scp, feats_scp)   # features.py:368-376: the scp now points at the final features
        for did in self.job.dictionary_ids:
            self.job.write_maps(did)


# ------------------------------------------------------------------------------------------------ a6: graphs
@dataclass
class CompileTrainGraphsArguments(MfaArguments):
    working_directory: Path
    lexicon_compilers: Dict[int, Lexicon]
    tree_path: Path
    model_path: Path
    use_g2p: bool = False


class CompileTrainGraphsFunction(KaldiFunction):
    def __init__(self, args: CompileTrainGraphsArguments):
        super().__init__(args)
        self.a = args

    def _run(self):
        for did, lexicon in self.a.lexicon_compilers.items():
            compiler = KC.TrainingGraphCompiler(self.a.model_path, self.a.tree_path, lexicon, use_g2p=self.a.use_g2p, batch_size=500)
            fst_ark = self.job.construct_path(self.a.working_directory, "fsts", "ark", did)
            records = [(u.kaldi_id, u.normalized_text) for u in self.job.utts(did)]
            compiler.export_graphs(fst_ark, records, callback=self.callback)
            del compiler


# ------------------------------------------------------------------------------------------------ a7: align
@dataclass
class AlignArguments(MfaArguments):
    working_directory: Path
    model_path: Path
    align_options: MetaDict
    confidence: bool = False
    final: bool = False
    silence_phone_ids: Sequence[int] = ()


class AlignFunction(KaldiFunction):
    def __init__(self, args: AlignArguments):
        super().__init__(args)
        self.a = args

    def _run(self):
        opts = dict(self.a.align_options)
        boost = opts.pop("boost_silence", 1.0)
        aligner = KC.GmmAligner(self.a.model_path, **opts)
        if boost != 1.0 and self.a.silence_phone_ids:
            aligner.boost_silence(boost, self.a.silence_phone_ids)
        first_pass = str(self.a.model_path).endswith(".alimdl")   # multiprocessing.py:841-863
        wd = Path(self.a.working_directory)
        for did in self.job.dictionary_ids:
            fst_path = self.job.construct_path(wd, "fsts", "ark", did)
            graphs = KC.FstArchive(fst_path)
            feats = self.job.construct_feature_archive(wd, did)
            ali = self.job.construct_path(wd, "ali", "ark", did)
            words = self.job.construct_path(wd, "words", "ark", did)
            likes = self.job.construct_path(wd, "likelihoods", "ark", did)
            # the reference removes the three targets before every export (alignment/multiprocessing.py:836-839): after a first pass
            # they are symlinks to the *_first_pass archives, and opening them for writing would clobber those
            links = (ali, words, likes)
            for path in links:
                Path(path).unlink(missing_ok=True)
            if first_pass:
                ali = self.job.construct_path(wd, "ali_first_pass", "ark", did)
                words = self.job.construct_path(wd, "words_first_pass", "ark", did)
                likes = self.job.construct_path(wd, "likelihoods_first_pass", "ark", did)
            aligner.export_alignments(ali, graphs, feats, word_file_name=words, likelihood_file_name=likes, callback=self.callback)
            if first_pass:
                for src, link in zip((ali, words, likes), links):
                    try:
                        os.symlink(Path(src).name, link)
                    except OSError:   # file systems without symlinks
                        shutil.copyfile(src, link)
            feats.close()


def align_utterances(jobs: Sequence[Job], working_directory, model_path, align_options: MetaDict, silence_phone_ids=(), training: bool = False,
                     num_threads: int = 1):
    """AlignMixin.align_utterances (alignment/mixins.py:282-380): drains (utt, loglike), stores per-utterance likelihood / frames,
    returns the workflow score (mean log-likelihood per utterance); raises NoAlignmentsError when nothing aligned."""
    args = [AlignArguments(j.id, j, None, Path(working_directory), Path(model_path), align_options, False, False, silence_phone_ids) for j in jobs]
    by_id = {u.kaldi_id: u for j in jobs for u in j.utterances}
    likes = []
    for utt_id, like in run_kaldi_function(AlignFunction, args, num_threads=num_threads):
        u = by_id[utt_id]
        u.alignment_log_likelihood = like
        likes.append(like)
    if not likes:
        raise NoAlignmentsError(len(by_id), align_options.get("beam", 10), align_options.get("retry_beam", 40))
    return float(np.mean(likes)), len(by_id) - len(likes)


# ------------------------------------------------------------------------------------------------ a9 / a10: statistics + update
@dataclass
class AccStatsArguments(MfaArguments):
    working_directory: Path
    model_path: Path


class AccStatsFunction(KaldiFunction):
    def __init__(self, args: AccStatsArguments):
        super().__init__(args)
        self.a = args

    def _run(self):
        wd = Path(self.a.working_directory)
        for did in self.job.dictionary_ids:
            acc = KC.GmmStatsAccumulator(self.a.model_path)
            feats = self.job.construct_feature_archive(wd, did)
            alis = KC.AlignmentArchive(self.job.construct_path(wd, "ali", "ark", did))
            acc.accumulate_stats(feats, alis, callback=self.callback)
            self.callback((acc.transition_accs, acc.gmm_accs))


def acc_stats(jobs: Sequence[Job], working_directory, iteration: int, mixup: int = 0, power: float = 0.25, min_gaussian_occupancy: float = 10.0,
              all_reduce: Optional[Callable[[np.ndarray], np.ndarray]] = None):
    """AcousticModelTrainingMixin.acc_stats (acoustic_modeling/base.py:277-338 upstream): sum the jobs' accumulators
    (`all_reduce`, when given, additionally sums across ranks: NCCL over the packed f64 block), MLE update, write (it+1).mdl.
    Returns (avg log-likelihood per frame, objective improvement, frames)."""
    wd = Path(working_directory)
    model_path = wd / f"{iteration}.mdl"
    tm, am = K.read_gmm_model(model_path)
    trans = tm.InitStats()
    gacc = AccumAmDiagGmm.init(am)
    args = [AccStatsArguments(j.id, j, None, wd, model_path) for j in jobs]
    for t, g in run_kaldi_function(AccStatsFunction, args):
        trans += t            # transition_accs.AddVec(1.0, t)
        gacc.Add(1.0, g)      # gmm_accs.Add(1.0, g)
    if all_reduce is not None:
        flat = np.concatenate([gacc.occ, gacc.mean.ravel(), gacc.var.ravel(), trans, [gacc.tot_like, gacc.tot_frames]])
        flat = all_reduce(flat)
        G, D = am.NumGauss(), am.dim
        gacc.occ = flat[:G]; gacc.mean = flat[G:G + G * D].reshape(G, D); gacc.var = flat[G + G * D:G + 2 * G * D].reshape(G, D)
        trans = flat[G + 2 * G * D:G + 2 * G * D + tm.num_tids + 1]
        gacc.tot_like, gacc.tot_frames = float(flat[-2]), float(flat[-1])
    # tm.mle_update(transition_accs) + am.mle_update(gmm_accs, mixup=, power=) of the reference: one device call (csrc/mstep.cu)
    new_am, impr, count = mle_update(am, gacc, mixup=mixup, power=power, min_gaussian_occupancy=min_gaussian_occupancy, tm=tm, transition_accs=trans)
    K.write_gmm_model(wd / f"{iteration + 1}.mdl", tm, new_am)
    avg = gacc.tot_like / max(gacc.tot_frames, 1.0)
    return avg, impr, gacc.tot_frames


# ------------------------------------------------------------------------------------------------ a11: equal alignment (mono iteration 0)
@dataclass
class MonoAlignEqualArguments(MfaArguments):
    working_directory: Path
    model_path: Path


class MonoAlignEqualFunction(KaldiFunction):
    """acoustic_modeling/monophone.py:40-139: equal alignment of every utterance through its training graph (Kaldi
    align-equal-compiled), ali.D.J.ark written, statistics of the flat-start model accumulated on those alignments.  Progress
    callbacks are utterance ids; the payload is (transition_accs, gmm_accs)."""

    def __init__(self, args: MonoAlignEqualArguments):
        super().__init__(args)
        self.a = args

    def _run(self):
        wd = Path(self.a.working_directory)
        for did in self.job.dictionary_ids:
            acc = KC.GmmStatsAccumulator(self.a.model_path)
            feat_path = self.job.construct_path(self.job.split_directory, "feats", "scp", did)
            if not feat_path.exists():
                feat_path = self.job.construct_path(self.job.split_directory, "feats", "scp")
            feats = KC.FeatureArchive(feat_path, deltas=True)   # monophone.py:96-99: deltas only, CMVN already applied (a4)
            graphs = KC.FstArchive(self.job.construct_path(wd, "fsts", "ark", did))
            keys = [k for k in graphs.keys() if k in set(feats.keys)]
            with K.ArkWriter(self.job.construct_path(wd, "ali", "ark", did)) as w:
                B = 512
                for i in range(0, len(keys), B):
                    ks = keys[i:i + B]
                    x, fo = feats.batch(ks)
                    fsts = [graphs[k] for k in ks]
                    res = KC.gmm_align_equal_batch(ks, fsts, np.diff(fo))
                    ali = np.zeros(int(fo[-1]), np.int32)
                    for j, (k, r) in enumerate(zip(ks, res)):
                        if r is None or len(r[0]) == 0:   # monophone.py:100-113: zero-length / empty graph / failed -> error, skipped
                            continue
                        ali[fo[j]:fo[j + 1]] = r[0]
                        w.write_int_vector(k, np.asarray(r[0], np.int32))
                        self.callback(k)
                    acc.accumulate_batch(x, ali)   # frames of skipped utterances carry tid 0 and are ignored by K4
            acc.sync()
            self.callback((acc.transition_accs, acc.gmm_accs))


def mono_align_equal(jobs: Sequence[Job], working_directory, all_reduce: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                     mixup: int = 0, power: float = 0.25):
    """MonophoneTrainer.mono_align_equal (acoustic_modeling/monophone.py:237-296): equal alignments + statistics with 0.mdl, then the
    first MLE update (min_gaussian_occupancy 3, mix-up to `mixup` = current_gaussians) -> 1.mdl.
    Returns (avg log-likelihood per frame, frames)."""
    wd = Path(working_directory)
    model_path = wd / "0.mdl"
    tm, am = K.read_gmm_model(model_path)
    trans = tm.InitStats()
    gacc = AccumAmDiagGmm.init(am)
    args = [MonoAlignEqualArguments(j.id, j, None, wd, model_path) for j in jobs]
    for r in run_kaldi_function(MonoAlignEqualFunction, args):
        if isinstance(r, tuple):
            trans += r[0]
            gacc.Add(1.0, r[1])
    if all_reduce is not None:
        flat = all_reduce(np.concatenate([gacc.occ, gacc.mean.ravel(), gacc.var.ravel(), trans, [gacc.tot_like, gacc.tot_frames]]))
        G, D = am.NumGauss(), am.dim
        gacc.occ = flat[:G]; gacc.mean = flat[G:G + G * D].reshape(G, D); gacc.var = flat[G + G * D:G + 2 * G * D].reshape(G, D)
        trans = flat[G + 2 * G * D:G + 2 * G * D + tm.num_tids + 1]
        gacc.tot_like, gacc.tot_frames = float(flat[-2]), float(flat[-1])
    new_am, _impr, _count = mle_update(am, gacc, mixup=mixup, power=power, min_gaussian_occupancy=3.0, tm=tm, transition_accs=trans)   # monophone.py:279-284
    K.write_gmm_model(wd / "1.mdl", tm, new_am)
    return gacc.tot_like / max(gacc.tot_frames, 1.0), gacc.tot_frames


# ------------------------------------------------------------------------------------------------ N2: fMLLR between the passes
@dataclass
class CalcFmllrArguments(MfaArguments):
    working_directory: Path
    ali_model_path: Path
    model_path: Path
    fmllr_options: MetaDict
    silence_phone_ids: Sequence[int] = ()


class CalcFmllrFunction(KaldiFunction):
    """corpus/features.py:423-548: per dictionary, speaker transforms from the job's alignments (two-model when ali_model_path is
    the .alimdl), composed with previous transforms when trans.D.J.scp exists, written to trans.D.J.ark/.scp in the split directory
    (where Job.construct_feature_archive picks them up for the second pass).  callback: (speaker, objf improvement, count)."""

    def __init__(self, args: CalcFmllrArguments):
        super().__init__(args)
        self.a = args

    def _run(self):
        wd = Path(self.a.working_directory)
        for did in self.job.dictionary_ids:
            ali_path = self.job.construct_path(wd, "ali", "ark", did)
            if not ali_path.exists():
                continue
            trans_scp = self.job.construct_path(self.job.split_directory, "trans", "scp", did)
            previous = KC.MatrixArchive(trans_scp) if trans_scp.exists() else None
            spk2utt: Dict[str, List[str]] = {}
            for u in self.job.utts(did):
                spk2utt.setdefault(str(u.speaker_id), []).append(u.kaldi_id)
            feats = self.job.construct_feature_archive(wd, did)
            computer = KC.FmllrComputer(self.a.ali_model_path, self.a.model_path, list(self.a.silence_phone_ids), spk2utt=spk2utt,
                                        **self.a.fmllr_options)
            alis = KC.AlignmentArchive(ali_path)
            tmp = self.job.construct_path(wd, "trans", "ark", did)
            out = computer.export_transforms(tmp, feats, alis, previous_transform_archive=previous, callback=self.callback)
            feats.close()
            final_ark = self.job.construct_path(self.job.split_directory, "trans", "ark", did)
            with K.ArkWriter(final_ark, trans_scp) as w:   # features.py:536-546: rewritten with an scp next to it
                for s, m in out.items():
                    w.write_matrix(str(s), np.asarray(m, np.float32))
            tmp.unlink()


def calc_fmllr(jobs: Sequence[Job], working_directory, ali_model_path, model_path, fmllr_options: Optional[MetaDict] = None,
               silence_phone_ids: Sequence[int] = ()):
    """AcousticCorpusMixin.calc_fmllr (corpus/acoustic_corpus.py:1370-1419).  Returns {speaker: (improvement, count)}."""
    opts = dict(fmllr_update_type="full", silence_weight=0.0, acoustic_scale=0.1)
    opts.update(fmllr_options or {})
    args = [CalcFmllrArguments(j.id, j, None, Path(working_directory), Path(ali_model_path), Path(model_path), opts, silence_phone_ids) for j in jobs]
    return {s: (impr, count) for s, impr, count in run_kaldi_function(CalcFmllrFunction, args)}


# ------------------------------------------------------------------------------------------------ N4: ali -> CTM -> TextGrid
@dataclass
class AlignmentExtractionArguments(MfaArguments):
    working_directory: Path
    model_path: Path
    lexicon_compilers: Dict[int, Lexicon]
    frame_shift: float = 0.01
    cleanup_textgrids: bool = True


class AlignmentExtractionFunction(KaldiFunction):
    """alignment/multiprocessing.py:1549-1862 (non-transcription branch): per dictionary, read ali/words/likelihoods archives, turn
    every alignment into word + phone intervals; callback payload (utterance id, dictionary id, HierarchicalCtm)."""

    def __init__(self, args: AlignmentExtractionArguments):
        super().__init__(args)
        self.a = args

    def _run(self):
        from . import export as X
        tm, _ = K.read_gmm_model(self.a.model_path)
        wd = Path(self.a.working_directory)
        for did in self.job.dictionary_ids:
            ali_path = self.job.construct_path(wd, "ali", "ark", did)
            if not ali_path.exists():
                continue
            lex = self.a.lexicon_compilers[did]
            by_key = {u.kaldi_id: u for u in self.job.utts(did)}
            archive = KC.AlignmentArchive(ali_path, self.job.construct_path(wd, "words", "ark", did), self.job.construct_path(wd, "likelihoods", "ark", did))
            for alignment in archive:
                u = by_key.get(alignment.utterance_id)
                if u is None:
                    continue
                ctm = X.alignment_to_ctm(alignment, tm, lex, self.a.frame_shift, u.begin, u.end, u.normalized_text or None)
                self.callback((u.id, did, ctm))


def export_textgrids(jobs: Sequence[Job], working_directory, model_path, lexicon_compilers: Dict[int, Lexicon], output_directory,
                     frame_shift: float = 0.01, output_format: str = "long_textgrid", cleanup_textgrids: bool = True) -> Dict[int, Path]:
    """CorpusAligner.collect_alignments + export_files (alignment/base.py:1549-1720, 2536-2748) without the database: one output file per
    sound file (utterances of the same `path` share it; several speakers -> "speaker - words" tiers).  Returns {utterance id: path}."""
    from . import export as X
    out_dir = Path(output_directory)
    out_dir.mkdir(parents=True, exist_ok=True)
    args = [AlignmentExtractionArguments(j.id, j, None, Path(working_directory), Path(model_path), lexicon_compilers, frame_shift, cleanup_textgrids)
            for j in jobs]
    by_id = {u.id: u for j in jobs for u in j.utterances}
    files: Dict[str, Dict[str, Dict[str, list]]] = {}
    durations: Dict[str, float] = {}
    written: Dict[int, Path] = {}
    for uid, did, ctm in run_kaldi_function(AlignmentExtractionFunction, args):
        u = by_id[uid]
        data = X.ctm_to_speaker_data(ctm, str(u.speaker_id), lexicon_compilers[did], cleanup_textgrids)
        spk = files.setdefault(u.path, {}).setdefault(str(u.speaker_id), {"words": [], "phones": []})
        for kind in ("words", "phones"):
            spk[kind].extend(data[str(u.speaker_id)][kind])
        end = u.end if u.end is not None else (u.begin or 0.0) + u.duration
        durations[u.path] = max(durations.get(u.path, 0.0), float(end))
        written[uid] = None
    ext = {"long_textgrid": ".TextGrid", "short_textgrid": ".TextGrid", "json": ".json", "csv": ".csv"}[output_format]
    for path, speaker_data in files.items():
        out = out_dir / (Path(path).stem + ext)
        X.export_textgrid(speaker_data, out, durations[path], frame_shift, output_format)
        for uid, u in by_id.items():
            if u.path == path and uid in written:
                written[uid] = out
    return written


# ------------------------------------------------------------------------------------------------ online path (section 3.2)
class AlignerError(Exception):
    """online/alignment.py:108-112: raised when the single utterance cannot be aligned."""


def align_utterance_online(model_path, tree_path, lexicon: Lexicon, pcm: np.ndarray, text: str, mfcc_options: Optional[dict] = None,
                           cmvn: Optional[np.ndarray] = None, lda_mat: Optional[np.ndarray] = None, fmllr_trans: Optional[np.ndarray] = None,
                           beam: float = 10, retry_beam: float = 40, transition_scale: float = 1.0, acoustic_scale: float = 0.1,
                           self_loop_scale: float = 0.1, boost_silence: float = 1.0, silence_phone_ids: Sequence[int] = (),
                           phone_table: Optional[Dict[int, str]] = None):
    """online/alignment.py:29-123 (align_utterance_online) without files: MFCC -> CMVN -> deltas | splice+LDA -> fMLLR ->
    compile_fst(text) -> GmmAligner.align_utterance -> phone CTM.  Returns (Alignment, [CtmInterval])."""
    mc = KC.MfccComputer(**(mfcc_options or {}))
    raw = mc.compute_mfccs(np.asarray(pcm, np.int16))
    if cmvn is None:
        cmvn = KC.CmvnComputer().compute_cmvn_from_features([raw])
    fo = np.asarray([0, raw.shape[0]], np.int64)
    fm = None if fmllr_trans is None else np.asarray(fmllr_trans, np.float32)[None]
    feats = KC.get_engine().features(raw, fo, "lda" if lda_mat is not None else "deltas", lda=lda_mat, fmllr=fm, cmvn_stats=np.asarray(cmvn)[None],
                                     utt2spk=np.zeros(1, np.int32), n_spk=1)
    compiler = KC.TrainingGraphCompiler(model_path, tree_path, lexicon)
    fst = compiler.compile_fst(text)
    aligner = KC.GmmAligner(model_path, transition_scale=transition_scale, acoustic_scale=acoustic_scale, self_loop_scale=self_loop_scale, beam=beam,
                            retry_beam=retry_beam)
    if boost_silence != 1.0 and silence_phone_ids:
        aligner.boost_silence(boost_silence, silence_phone_ids)
    ali = aligner.align_utterance(fst, feats)
    if ali is None:
        raise AlignerError(f"Could not align the file with the current beam size ({beam}, please try increasing the beam size via `--beam X`")
    return ali, ali.generate_ctm(aligner.transition_model, phone_table or {}, mc.frame_shift / 1000.0)


def align_utterance_online_ctm(model_path, tree_path, lexicon: Lexicon, utterance: "KC.KalpyUtterance", mfcc_computer: Optional[KC.MfccComputer] = None,
                               cmvn: Optional[np.ndarray] = None, lda_mat: Optional[np.ndarray] = None, fmllr_trans: Optional[np.ndarray] = None,
                               uses_cmvn: bool = True, beam: float = 10, retry_beam: float = 40, transition_scale: float = 1.0,
                               acoustic_scale: float = 0.1, self_loop_scale: float = 0.1, boost_silence: float = 1.0,
                               silence_phone_ids: Sequence[int] = ()):
    """online/alignment.py:29-123 with the reference's own argument shape: a KalpyUtterance in, a HierarchicalCtm (word intervals owning
    their phone intervals, shifted to the segment's position in the file) out.  AlignerError when the beam is too tight."""
    from . import export as X
    mc = mfcc_computer or KC.MfccComputer()
    if utterance.mfccs is None:
        utterance.generate_mfccs(mc)
        if uses_cmvn:
            if cmvn is None:
                cmvn = KC.CmvnComputer().compute_cmvn_from_features([utterance.mfccs])
            utterance.apply_cmvn(cmvn)
    feats = utterance.generate_features(mc, None, lda_mat=lda_mat, fmllr_trans=fmllr_trans)
    fst = KC.TrainingGraphCompiler(model_path, tree_path, lexicon).compile_fst(utterance.transcript)
    aligner = KC.GmmAligner(model_path, transition_scale=transition_scale, acoustic_scale=acoustic_scale, self_loop_scale=self_loop_scale,
                            beam=beam, retry_beam=retry_beam)
    if boost_silence != 1.0 and silence_phone_ids:
        aligner.boost_silence(boost_silence, silence_phone_ids)
    alignment = aligner.align_utterance(fst, feats)
    if alignment is None:
        raise AlignerError(f"Could not align the file with the current beam size ({beam}, please try increasing the beam size via `--beam X`")
    return X.alignment_to_ctm(alignment, aligner.transition_model, lexicon, mc.frame_shift / 1000.0, utterance.segment.begin, utterance.segment.end,
                              utterance.transcript)


def align_one(sound_file_path, utterances: Sequence[Tuple[Optional[float], Optional[float], int, str]], model_path, tree_path, lexicon: Lexicon,
              output_path, file_duration: float, lda_mat: Optional[np.ndarray] = None, output_format: str = "long_textgrid", **align_options):
    """`mfa align_one` (command_line/align_one.py:157-196) without the CLI: every (begin, end, channel, text) segment of one sound file
    -> MFCCs, ONE CMVN over the file's segments, per-segment online alignment, one TextGrid.  Returns the merged HierarchicalCtm."""
    from . import export as X
    mc = KC.MfccComputer()
    utts = []
    for begin, end, channel, text in utterances:
        u = KC.KalpyUtterance(KC.Segment(str(sound_file_path), begin, end, channel), text)
        u.generate_mfccs(mc)
        utts.append(u)
    cmvn = KC.CmvnComputer().compute_cmvn_from_features([u.mfccs for u in utts])
    file_ctm = X.HierarchicalCtm([])
    for u in utts:
        u.apply_cmvn(cmvn)
        ctm = align_utterance_online_ctm(model_path, tree_path, lexicon, u, mc, lda_mat=lda_mat, **align_options)
        file_ctm.word_intervals.extend(ctm.word_intervals)
    if str(output_path) != "-":
        Path(output_path).parent.mkdir(parents=True, exist_ok=True)
        file_ctm.export_textgrid(output_path, file_duration=file_duration, output_format=output_format, silence_word=lexicon.silence_word)
    return file_ctm


def two_pass_align_pcm(engine, ali_model, model, graphs, pcm, sample_off, utt2spk, n_spk: int, mfcc_opts, feat_mode: str = "lda", lda=None,
                       silence_phone_ids: Sequence[int] = (), silence_weight: float = 0.0, align_opts=None, workspace_bytes: int = 0,
                       min_count: float = 500.0, num_iters: int = 40):
    """CorpusAligner.align's two passes (alignment/base.py:491-558) for one in-memory batch, nothing leaving the GPU but the per-speaker
    transforms: pass 1 with the speaker-independent model -> per-speaker fMLLR statistics and transforms (K5) -> pass 2 with the
    speaker-adapted model on the transformed features.  `ali_model` / `model` are engine.DeviceModel (the same object when there is no
    separate .alimdl), `graphs` an engine.Graphs, `pcm` a numpy array or a torch cuda tensor (then every intermediate stays on the device).
    Returns (pass-1 AlignResult, transforms [n_spk, D, D+1] float32 numpy, (objf improvement, count) per speaker, pass-2 AlignResult)."""
    from . import engine as E
    r1 = E.align_pcm(engine, ali_model, graphs, pcm, sample_off, utt2spk, n_spk, mfcc_opts, feat_mode, lda=lda, align=align_opts,
                     workspace_bytes=workspace_bytes)
    raw, fo = engine.mfcc(pcm, sample_off, mfcc_opts)
    stats = engine.cmvn_stats(raw, fo, utt2spk, n_spk)
    engine.sync()
    cm = stats.cpu().numpy() if E._is_torch(stats) else stats
    feats = engine.features(raw, fo, feat_mode, lda=lda, cmvn_stats=cm, utt2spk=utt2spk, n_spk=n_spk)
    tm = model.tm
    sil = np.isin(tm.tid2phone, np.asarray(sorted(int(p) for p in silence_phone_ids), tm.tid2phone.dtype))
    tw = np.where(sil, np.float32(silence_weight), np.float32(1.0)).astype(np.float32)
    tw[0] = 0.0
    T = int(fo[-1])
    ali = r1.ali[:T] if E._is_torch(r1.ali) else np.ascontiguousarray(r1.ali[:T])
    fstats = model.fmllr_acc(feats, ali.contiguous() if E._is_torch(ali) else ali, fo, utt2spk, n_spk, tid_weight=tw,
                             post_model=ali_model if ali_model is not model else None)
    W, impr, count = engine.fmllr_update(fstats, model.dim, num_iters, min_count)
    engine.sync()   # `raw` / `feats` / `fstats` may be recycled by torch from here on
    Wh = W.cpu().numpy() if E._is_torch(W) else np.asarray(W)
    r2 = E.align_pcm(engine, model, graphs, pcm, sample_off, utt2spk, n_spk, mfcc_opts, feat_mode, lda=lda, fmllr=Wh, align=align_opts,
                     workspace_bytes=workspace_bytes)
    return r1, Wh, (impr, count), r2
