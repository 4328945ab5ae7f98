"""Kaldi binary object / table formats used on MFA's alignment hot path.

Host-side (numpy) readers and writers for the files that cross the reference's
stage boundaries: ``final.mdl`` (TransitionModel +
AmDiagGmm), ``tree`` (ContextDependency EventMap), ``lda.mat``, and the ark/scp
tables ``feats``/``cmvn``/``ali``/``words``/``likelihoods``/``trans``/``fsts``.

Reference call sites these replace (kalpy is not vendored in /root/reference):
  * ``read_gmm_model``        montreal_forced_aligner/alignment/multiprocessing.py:1393
  * ``read_kaldi_object``     montreal_forced_aligner/alignment/multiprocessing.py:1218
  * ``CompressedMatrixWriter``montreal_forced_aligner/corpus/features.py:209
  * ``Int32VectorWriter``     montreal_forced_aligner/acoustic_modeling/monophone.py:95
"""
from __future__ import annotations

import io
import math
import os
import struct
from dataclasses import dataclass, field
from typing import BinaryIO, Dict, Iterator, List, Optional, Tuple

import numpy as np

M_LOG_2PI = 1.8378770664093454835606594728112


# --------------------------------------------------------------------------- low level
class _Reader:
    def __init__(self, f: BinaryIO):
        self.f = f

    def peek(self, n: int = 1) -> bytes:
        pos = self.f.tell()
        b = self.f.read(n)
        self.f.seek(pos)
        return b

    def token(self) -> str:
        out = bytearray()
        while True:
            c = self.f.read(1)
            if not c:
                break
            if c in b" \n\t":
                if out:
                    break
                continue
            out += c
        return out.decode("utf8")

    def expect(self, tok: str):
        t = self.token()
        if t != tok:
            raise ValueError(f"expected token {tok!r}, got {t!r} at {self.f.tell()}")

    def int32(self) -> int:
        m = self.f.read(1)[0]
        if m == 4:
            return struct.unpack("<i", self.f.read(4))[0]
        if m == 0xFC:  # -4: unsigned 32-bit
            return struct.unpack("<I", self.f.read(4))[0]
        if m == 8:
            return struct.unpack("<q", self.f.read(8))[0]
        if m == 1:
            return struct.unpack("<b", self.f.read(1))[0]
        raise ValueError(f"bad int marker {m} at {self.f.tell()}")

    def float32(self) -> float:
        m = self.f.read(1)[0]
        if m == 4:
            return struct.unpack("<f", self.f.read(4))[0]
        if m == 8:
            return struct.unpack("<d", self.f.read(8))[0]
        raise ValueError(f"bad float marker {m}")

    def int_vector(self) -> np.ndarray:
        esz = self.f.read(1)[0]
        if esz != 4:
            raise ValueError(f"int vector element size {esz}")
        n = struct.unpack("<i", self.f.read(4))[0]
        return np.frombuffer(self.f.read(4 * n), dtype="<i4").copy()

    def vector(self) -> np.ndarray:
        t = self.token()
        if t not in ("FV", "DV"):
            raise ValueError(f"expected FV/DV got {t!r}")
        n = self.int32()
        dt, w = ("<f4", 4) if t == "FV" else ("<f8", 8)
        return np.frombuffer(self.f.read(w * n), dtype=dt).copy()

    def matrix(self) -> np.ndarray:
        t = self.token()
        if t in ("CM", "CM2", "CM3"):
            return _read_compressed_body(self.f, t)
        if t not in ("FM", "DM"):
            raise ValueError(f"expected FM/DM got {t!r}")
        r = self.int32()
        c = self.int32()
        dt, w = ("<f4", 4) if t == "FM" else ("<f8", 8)
        return np.frombuffer(self.f.read(w * r * c), dtype=dt).reshape(r, c).copy()


def _w_token(f, tok: str):
    f.write(tok.encode("utf8") + b" ")


def _w_int32(f, v: int):
    f.write(b"\x04" + struct.pack("<i", int(v)))


def _w_float(f, v: float):
    f.write(b"\x04" + struct.pack("<f", float(v)))


def write_vector(f, v: np.ndarray):
    v = np.asarray(v)
    if v.dtype == np.float64:
        _w_token(f, "DV")
        _w_int32(f, v.shape[0])
        f.write(v.astype("<f8").tobytes())
    else:
        _w_token(f, "FV")
        _w_int32(f, v.shape[0])
        f.write(v.astype("<f4").tobytes())


def write_matrix(f, m: np.ndarray):
    m = np.asarray(m)
    if m.dtype == np.float64:
        _w_token(f, "DM")
        dt = "<f8"
    else:
        _w_token(f, "FM")
        dt = "<f4"
    _w_int32(f, m.shape[0])
    _w_int32(f, m.shape[1])
    f.write(np.ascontiguousarray(m, dtype=dt).tobytes())


def write_int_vector(f, v: np.ndarray):
    v = np.asarray(v, dtype="<i4")
    f.write(b"\x04" + struct.pack("<i", v.shape[0]) + v.tobytes())


# --------------------------------------------------------------------------- CompressedMatrix
# Kaldi matrix/compressed-matrix.cc.  Format kOneByteWithColHeaders ("CM"):
# global header {min, range, rows, cols}, per column 4 x uint16 percentiles, then column-major
# uint8.  "CM2" = two-byte ints, "CM3" = one byte without column headers.


def _u16_to_float(minv, rng, u):
    return minv + rng * 1.52590218966964e-05 * u.astype(np.float32)


def _read_compressed_body(f, tok: str) -> np.ndarray:
    minv, rng, rows, cols = struct.unpack("<ffii", f.read(16))
    minv = np.float32(minv)
    rng = np.float32(rng)
    if tok == "CM":
        hdr = np.frombuffer(f.read(8 * cols), dtype="<u2").reshape(cols, 4)
        data = np.frombuffer(f.read(rows * cols), dtype=np.uint8).reshape(cols, rows)
        p0 = _u16_to_float(minv, rng, hdr[:, 0])[:, None]
        p25 = _u16_to_float(minv, rng, hdr[:, 1])[:, None]
        p75 = _u16_to_float(minv, rng, hdr[:, 2])[:, None]
        p100 = _u16_to_float(minv, rng, hdr[:, 3])[:, None]
        v = data.astype(np.float32)
        out = np.where(
            data <= 64,
            p0 + (p25 - p0) * v * np.float32(1 / 64.0),
            np.where(
                data <= 192,
                p25 + (p75 - p25) * (v - 64) * np.float32(1 / 128.0),
                p75 + (p100 - p75) * (v - 192) * np.float32(1 / 63.0),
            ),
        ).astype(np.float32)
        return np.ascontiguousarray(out.T)
    if tok == "CM2":
        data = np.frombuffer(f.read(2 * rows * cols), dtype="<u2").reshape(rows, cols)
        return (minv + rng * np.float32(1.0 / 65535.0) * data.astype(np.float32)).astype(np.float32)
    data = np.frombuffer(f.read(rows * cols), dtype=np.uint8).reshape(rows, cols)
    return (minv + rng * np.float32(1.0 / 255.0) * data.astype(np.float32)).astype(np.float32)


def compress_matrix(mat: np.ndarray) -> bytes:
    """Encode ``mat`` the way Kaldi's ``CompressedMatrix(mat)`` (kAutomaticMethod) does.

    rows > 8 -> "CM" (one byte + column headers); else "CM2" (two-byte).  Returns the object
    bytes starting at the token (caller adds the ``\\0B`` binary marker).
    """
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    rows, cols = mat.shape
    out = io.BytesIO()
    if rows == 0 or cols == 0:
        _w_token(out, "CM")
        out.write(struct.pack("<ffii", 0.0, 0.0, 0, 0))
        return out.getvalue()
    minv = np.float32(mat.min())
    maxv = np.float32(mat.max())
    if maxv == minv:
        maxv = np.float32(minv + (1.0 + abs(float(minv))))
    rng = np.float32(maxv - minv)
    if rows > 8:
        _w_token(out, "CM")
        out.write(struct.pack("<ffii", float(minv), float(rng), rows, cols))

        def f2u16(v):
            f = (v.astype(np.float32) - minv) / rng
            f = np.clip(f, 0.0, 1.0)
            return (f * np.float32(65535) + np.float32(0.499)).astype(np.int64)

        srt = np.sort(mat, axis=0)
        if rows >= 5:
            q = rows // 4
            p0, p25, p75, p100 = srt[0], srt[q], srt[3 * q], srt[rows - 1]
        else:  # pragma: no cover (rows>8 here)
            p0 = srt[0]
            p25 = srt[1] if rows > 1 else p0 + 1
            p75 = srt[2] if rows > 2 else p25 + 1
            p100 = srt[3] if rows > 3 else p75 + 1
        u0 = np.minimum(f2u16(p0), 65532)
        u25 = np.minimum(np.maximum(f2u16(p25), u0 + 1), 65533)
        u75 = np.minimum(np.maximum(f2u16(p75), u25 + 1), 65534)
        u100 = np.maximum(f2u16(p100), u75 + 1)
        hdr = np.stack([u0, u25, u75, u100], axis=1).astype("<u2")
        out.write(hdr.tobytes())
        f0 = _u16_to_float(minv, rng, hdr[:, 0])[None, :]
        f25 = _u16_to_float(minv, rng, hdr[:, 1])[None, :]
        f75 = _u16_to_float(minv, rng, hdr[:, 2])[None, :]
        f100 = _u16_to_float(minv, rng, hdr[:, 3])[None, :]
        v = mat
        lo = ((v - f0) / (f25 - f0) * np.float32(64) + np.float32(0.5)).astype(np.int64)
        mid = 64 + ((v - f25) / (f75 - f25) * np.float32(128) + np.float32(0.5)).astype(np.int64)
        hi = 192 + ((v - f75) / (f100 - f75) * np.float32(63) + np.float32(0.5)).astype(np.int64)
        lo = np.clip(lo, 0, 64)
        mid = np.clip(mid, 64, 192)
        hi = np.clip(hi, 192, 255)
        q8 = np.where(v < f25, lo, np.where(v < f75, mid, hi)).astype(np.uint8)
        out.write(np.ascontiguousarray(q8.T).tobytes())
    else:
        _w_token(out, "CM2")
        out.write(struct.pack("<ffii", float(minv), float(rng), rows, cols))
        f = (mat - minv) / rng
        u = (np.clip(f, 0, 1) * np.float32(65535) + np.float32(0.499)).astype("<u2")
        out.write(u.tobytes())
    return out.getvalue()


def decompress_matrix(blob: bytes) -> np.ndarray:
    f = io.BytesIO(blob)
    return _Reader(f).matrix()


# --------------------------------------------------------------------------- model objects
@dataclass
class HmmState:
    forward_pdf_class: int
    self_loop_pdf_class: int
    transitions: List[Tuple[int, float]]  # (dest hmm state, prob)


@dataclass
class Topology:
    phones: np.ndarray  # sorted phone ids
    phone2idx: np.ndarray  # phone id -> entry index (-1 if none)
    entries: List[List[HmmState]]

    def states_for(self, phone: int) -> List[HmmState]:
        return self.entries[int(self.phone2idx[phone])]


class TransitionModel:
    """Restatement of Kaldi ``hmm/transition-model.{h,cc}`` bookkeeping."""

    def __init__(self, topo: Topology, tuples: np.ndarray, log_probs: np.ndarray):
        self.topo = topo
        self.tuples = tuples  # [n_tstates, 4] (phone, hmm_state, fwd_pdf, self_loop_pdf); tstate ids are 1-based
        self.log_probs = log_probs.astype(np.float32)  # [n_tids + 1], index 0 unused
        self._derive()

    def _derive(self):
        nts = self.tuples.shape[0]
        state2id = np.zeros(nts + 2, dtype=np.int32)  # tstate -> first tid
        cur = 1
        for ts in range(1, nts + 1):
            state2id[ts] = cur
            ph, hs = int(self.tuples[ts - 1, 0]), int(self.tuples[ts - 1, 1])
            cur += len(self.topo.states_for(ph)[hs].transitions)
        state2id[nts + 1] = cur
        self.state2id = state2id
        ntid = cur - 1
        self.num_tids = ntid
        if self.log_probs.shape[0] != ntid + 1:
            raise ValueError(f"log_probs size {self.log_probs.shape[0]} != num_tids+1 {ntid + 1}")
        id2state = np.zeros(ntid + 1, dtype=np.int32)
        tid2pdf = np.full(ntid + 1, -1, dtype=np.int32)
        tid2phone = np.zeros(ntid + 1, dtype=np.int32)
        is_self = np.zeros(ntid + 1, dtype=np.int8)
        is_final = np.zeros(ntid + 1, dtype=np.int8)
        self_loop_tid = np.zeros(nts + 1, dtype=np.int32)  # tstate -> its self-loop tid (0 if none)
        for ts in range(1, nts + 1):
            ph, hs, fpdf, spdf = (int(x) for x in self.tuples[ts - 1])
            states = self.topo.states_for(ph)
            nfinal = len(states) - 1
            for k, (dst, _p) in enumerate(states[hs].transitions):
                tid = state2id[ts] + k
                id2state[tid] = ts
                tid2phone[tid] = ph
                if dst == hs:
                    is_self[tid] = 1
                    tid2pdf[tid] = spdf
                    self_loop_tid[ts] = tid
                else:
                    tid2pdf[tid] = fpdf
                if dst == nfinal:
                    is_final[tid] = 1
        self.id2state = id2state
        self.tid2pdf = tid2pdf
        self.tid2phone = tid2phone
        self.is_self_loop = is_self
        self.is_final_tid = is_final
        self.self_loop_tid = self_loop_tid
        # non-self-loop log prob per tstate (ComputeDerivedOfProbs)
        nsl = np.zeros(nts + 1, dtype=np.float32)
        for ts in range(1, nts + 1):
            sl = self_loop_tid[ts]
            if sl != 0:
                p = 1.0 - math.exp(float(self.log_probs[sl]))
                if p <= 0.0:
                    p = 1.0e-10
                nsl[ts] = np.float32(math.log(p))
        self.non_self_loop_log_prob = nsl
        self.num_pdfs_in_tm = int(max(self.tuples[:, 2].max(), self.tuples[:, 3].max())) + 1
        # (phone, hmm_state, fwd_pdf, self_loop_pdf) -> tstate
        self._tuple2state = {tuple(int(x) for x in row): i + 1 for i, row in enumerate(self.tuples)}

    # --- kalpy/Kaldi-named accessors used by MFA
    def NumTransitionIds(self) -> int:
        return self.num_tids

    def NumPdfs(self) -> int:
        return self.num_pdfs_in_tm

    def TransitionIdToPdf(self, tid: int) -> int:
        return int(self.tid2pdf[tid])

    def TransitionIdToPhone(self, tid: int) -> int:
        return int(self.tid2phone[tid])

    def tuple_to_tstate(self, phone: int, hmm_state: int, fpdf: int, spdf: int) -> int:
        return self._tuple2state[(phone, hmm_state, fpdf, spdf)]

    def scaled_transition_log_probs(self, transition_scale: float, self_loop_scale: float) -> np.ndarray:
        """Per-tid value added (negated) to graph arcs by ``AddTransitionProbs`` (hmm-utils.cc)."""
        lp = self.log_probs.astype(np.float32)
        if transition_scale == self_loop_scale:
            out = lp * np.float32(transition_scale)
        else:
            nsl = self.non_self_loop_log_prob[self.id2state]
            out = np.where(
                self.is_self_loop == 1,
                np.float32(self_loop_scale) * lp,
                np.float32(self_loop_scale) * nsl + np.float32(transition_scale) * (lp - nsl),
            ).astype(np.float32)
        out[0] = 0.0
        return out

    def acc_stats(self, alignment, stats: np.ndarray) -> np.ndarray:
        """TransitionModel::Accumulate over an alignment (acoustic_modeling/monophone.py:120): stats[tid] += 1."""
        np.add.at(stats, np.asarray(alignment, np.int64), 1.0)
        return stats

    def InitStats(self) -> np.ndarray:
        return np.zeros(self.num_tids + 1, dtype=np.float64)

    def mle_update(self, stats: np.ndarray, floor: float = 0.01, mincount: float = 5.0):
        """Kaldi ``TransitionModel::MleUpdate`` (non-shared). Returns (objf_impr, count)."""
        objf_impr = 0.0
        count_sum = 0.0
        new_lp = self.log_probs.copy()
        for ts in range(1, self.tuples.shape[0] + 1):
            a, b = int(self.state2id[ts]), int(self.state2id[ts + 1])
            n = b - a
            if n <= 1:
                continue
            counts = stats[a:b].astype(np.float64)
            tot = counts.sum()
            count_sum += tot
            if tot < mincount:
                continue
            old = np.exp(self.log_probs[a:b].astype(np.float64))
            new = counts.copy()
            for _ in range(3):   # transition-model.cc MleUpdate: renormalise, THEN floor, three times (the floor is the last step)
                new = new / new.sum()
                new = np.maximum(new, floor)
            for k in range(n):
                if counts[k] > 0 and old[k] > 0 and new[k] > 0:
                    objf_impr += counts[k] * (math.log(new[k]) - math.log(old[k]))
            new_lp[a:b] = np.log(new).astype(np.float32)
        self.log_probs = new_lp
        self._derive()
        return objf_impr, count_sum


class AmDiagGmm:
    """All pdfs of the acoustic model, packed (Gaussians of pdf j are rows offsets[j]:offsets[j+1])."""

    def __init__(self, dim, offsets, weights, means_invvars, inv_vars, gconsts=None):
        self.dim = int(dim)
        self.offsets = np.asarray(offsets, dtype=np.int32)
        self.weights = np.ascontiguousarray(weights, dtype=np.float32)
        self.means_invvars = np.ascontiguousarray(means_invvars, dtype=np.float32)
        self.inv_vars = np.ascontiguousarray(inv_vars, dtype=np.float32)
        self.stored_gconsts = None if gconsts is None else np.asarray(gconsts, dtype=np.float32)
        self.gconsts = self.compute_gconsts()

    def compute_gconsts(self) -> np.ndarray:
        """DiagGmm::ComputeGconsts — recomputed on read as Kaldi does."""
        d = self.dim
        iv = self.inv_vars.astype(np.float64)
        miv = self.means_invvars.astype(np.float64)
        offset = -0.5 * M_LOG_2PI * d
        with np.errstate(divide="ignore"):
            gc = np.log(self.weights.astype(np.float32)).astype(np.float64) + offset
        gc = gc + 0.5 * np.log(iv).sum(1) - 0.5 * (miv * miv / iv).sum(1)
        gc = np.where(np.isfinite(gc), gc, -1.0e20)  # Kaldi: NaN/inf gconst -> very negative
        return gc.astype(np.float32)

    def NumPdfs(self) -> int:
        return self.offsets.shape[0] - 1

    def NumGauss(self) -> int:
        return int(self.offsets[-1])

    def Dim(self) -> int:
        return self.dim

    def means(self) -> np.ndarray:
        return (self.means_invvars.astype(np.float64) / self.inv_vars.astype(np.float64))

    def variances(self) -> np.ndarray:
        return 1.0 / self.inv_vars.astype(np.float64)

    def copy(self) -> "AmDiagGmm":
        return AmDiagGmm(self.dim, self.offsets.copy(), self.weights.copy(), self.means_invvars.copy(),
                         self.inv_vars.copy())


def _read_topology(r: _Reader) -> Topology:
    r.expect("<Topology>")
    phones = r.int_vector()
    phone2idx = r.int_vector()
    n = r.int32()
    is_hmm = True
    if n == -1:
        is_hmm = False
        n = r.int32()
    entries = []
    for _ in range(n):
        ns = r.int32()
        states = []
        for _ in range(ns):
            fpc = r.int32()
            spc = fpc if is_hmm else r.int32()
            nt = r.int32()
            trans = []
            for _ in range(nt):
                dst = r.int32()
                p = r.float32()
                trans.append((dst, p))
            states.append(HmmState(fpc, spc, trans))
        entries.append(states)
    r.expect("</Topology>")
    return Topology(phones, phone2idx, entries)


def read_topology_text(path_or_text) -> Topology:
    """Kaldi's TEXT topology format (what MFA writes as ``topo``: reference tests/data/dictionaries/expected/topo; kalpy
    ``read_topology``): <TopologyEntry> <ForPhones> ids </ForPhones> <State> i [<PdfClass> c | <ForwardPdfClass> f <SelfLoopPdfClass> s]
    (<Transition> dst p)* </State> ... </TopologyEntry>."""
    text = path_or_text if "<Topology>" in str(path_or_text) else open(path_or_text, "r").read()
    tok = text.split()
    i = 0

    def expect(t):
        nonlocal i
        if tok[i] != t:
            raise ValueError(f"topology: expected {t}, got {tok[i]}")
        i += 1
    expect("<Topology>")
    entries, for_phones = [], []
    while tok[i] == "<TopologyEntry>":
        i += 1
        expect("<ForPhones>")
        ph = []
        while tok[i] != "</ForPhones>":
            ph.append(int(tok[i])); i += 1
        i += 1
        states = []
        while tok[i] == "<State>":
            i += 1
            idx = int(tok[i]); i += 1
            if idx != len(states):
                raise ValueError("topology: states must be numbered consecutively")
            fpc = spc = -1
            if tok[i] == "<PdfClass>":
                fpc = spc = int(tok[i + 1]); i += 2
            elif tok[i] == "<ForwardPdfClass>":
                fpc = int(tok[i + 1]); i += 2
                expect("<SelfLoopPdfClass>")
                spc = int(tok[i]); i += 1
            trans = []
            while tok[i] in ("<Transition>", "<Final>"):
                if tok[i] == "<Final>":      # old format: probability of leaving through the final state
                    trans.append((len(states) + 1, float(tok[i + 1]))); i += 2
                else:
                    trans.append((int(tok[i + 1]), float(tok[i + 2]))); i += 3
            expect("</State>")
            states.append(HmmState(fpc, spc, trans))
        expect("</TopologyEntry>")
        entries.append(states)
        for_phones.append(ph)
    expect("</Topology>")
    phones = np.asarray(sorted(p for ph in for_phones for p in ph), np.int32)
    phone2idx = np.full(int(phones.max()) + 1 if phones.size else 1, -1, np.int32)
    for e, ph in enumerate(for_phones):
        for p in ph:
            phone2idx[p] = e
    return Topology(phones, phone2idx, entries)


def _write_topology(f, topo: Topology):
    _w_token(f, "<Topology>")
    write_int_vector(f, topo.phones)
    write_int_vector(f, topo.phone2idx)
    is_hmm = all(s.forward_pdf_class == s.self_loop_pdf_class for e in topo.entries for s in e)
    if not is_hmm:
        _w_int32(f, -1)
    _w_int32(f, len(topo.entries))
    for e in topo.entries:
        _w_int32(f, len(e))
        for s in e:
            _w_int32(f, s.forward_pdf_class)
            if not is_hmm:
                _w_int32(f, s.self_loop_pdf_class)
            _w_int32(f, len(s.transitions))
            for dst, p in s.transitions:
                _w_int32(f, dst)
                _w_float(f, p)
    _w_token(f, "</Topology>")


def read_transition_model(r: _Reader) -> TransitionModel:
    r.expect("<TransitionModel>")
    topo = _read_topology(r)
    tok = r.token()
    n = r.int32()
    if tok == "<Triples>":
        t = np.zeros((n, 4), dtype=np.int32)
        for i in range(n):
            t[i, 0] = r.int32()
            t[i, 1] = r.int32()
            t[i, 2] = r.int32()
            t[i, 3] = t[i, 2]
        r.expect("</Triples>")
    elif tok == "<Tuples>":
        t = np.zeros((n, 4), dtype=np.int32)
        for i in range(n):
            for k in range(4):
                t[i, k] = r.int32()
        r.expect("</Tuples>")
    else:
        raise ValueError(f"unexpected {tok}")
    r.expect("<LogProbs>")
    lp = r.vector()
    r.expect("</LogProbs>")
    r.expect("</TransitionModel>")
    return TransitionModel(topo, t, lp)


def read_am_diag_gmm(r: _Reader) -> AmDiagGmm:
    r.expect("<DIMENSION>")
    dim = r.int32()
    r.expect("<NUMPDFS>")
    npdf = r.int32()
    offs = [0]
    gcs, ws, mivs, ivs = [], [], [], []
    for _ in range(npdf):
        tok = r.token()
        if tok == "<DiagGMMBegin>":
            tok = r.token()
        if tok != "<DiagGMM>":
            raise ValueError(f"expected <DiagGMM> got {tok}")
        tok = r.token()
        gc = None
        if tok == "<GCONSTS>":
            gc = r.vector()
            tok = r.token()
        if tok != "<WEIGHTS>":
            raise ValueError(tok)
        w = r.vector()
        r.expect("<MEANS_INVVARS>")
        miv = r.matrix()
        r.expect("<INV_VARS>")
        iv = r.matrix()
        tok = r.token()
        if tok not in ("</DiagGMM>", "<DiagGMMEnd>"):
            raise ValueError(tok)
        offs.append(offs[-1] + w.shape[0])
        gcs.append(gc if gc is not None else np.zeros_like(w))
        ws.append(w)
        mivs.append(miv)
        ivs.append(iv)
    return AmDiagGmm(dim, np.array(offs), np.concatenate(ws), np.concatenate(mivs), np.concatenate(ivs),
                     np.concatenate(gcs))


def read_gmm_model(path) -> Tuple[TransitionModel, AmDiagGmm]:
    with open(path, "rb") as f:
        if f.read(2) != b"\0B":
            raise ValueError(f"{path}: only Kaldi binary models are supported")
        r = _Reader(f)
        tm = read_transition_model(r)
        am = read_am_diag_gmm(r)
    return tm, am


def write_gmm_model(path, tm: TransitionModel, am: AmDiagGmm):
    with open(path, "wb") as f:
        f.write(b"\0B")
        _w_token(f, "<TransitionModel>")
        _write_topology(f, tm.topo)
        triples = bool(np.all(tm.tuples[:, 2] == tm.tuples[:, 3]))
        _w_token(f, "<Triples>" if triples else "<Tuples>")
        _w_int32(f, tm.tuples.shape[0])
        for row in tm.tuples:
            for k in range(3 if triples else 4):
                _w_int32(f, row[k])
        _w_token(f, "</Triples>" if triples else "</Tuples>")
        _w_token(f, "<LogProbs>")
        write_vector(f, tm.log_probs.astype(np.float32))
        _w_token(f, "</LogProbs>")
        _w_token(f, "</TransitionModel>")
        _w_token(f, "<DIMENSION>")
        _w_int32(f, am.dim)
        _w_token(f, "<NUMPDFS>")
        _w_int32(f, am.NumPdfs())
        gc = am.compute_gconsts()
        for j in range(am.NumPdfs()):
            a, b = am.offsets[j], am.offsets[j + 1]
            _w_token(f, "<DiagGMM>")
            _w_token(f, "<GCONSTS>")
            write_vector(f, gc[a:b])
            _w_token(f, "<WEIGHTS>")
            write_vector(f, am.weights[a:b])
            _w_token(f, "<MEANS_INVVARS>")
            write_matrix(f, am.means_invvars[a:b])
            _w_token(f, "<INV_VARS>")
            write_matrix(f, am.inv_vars[a:b])
            _w_token(f, "</DiagGMM>")


# --------------------------------------------------------------------------- tree
@dataclass
class ContextDependency:
    """Kaldi ``tree/context-dep.h`` + EventMap, flattened.

    nodes[i] = (type, key, a, b): type 0=CE (a=answer); 1=SE (a=yes child, b=no child, yes-set in
    ``sets[i]``); 2=TE (children in ``tables[i]``, -1 = NULL).
    """
    N: int
    P: int
    nodes: List[Tuple[int, int, int, int]] = field(default_factory=list)
    sets: Dict[int, np.ndarray] = field(default_factory=dict)
    tables: Dict[int, List[int]] = field(default_factory=dict)
    root: int = 0

    def lookup(self, context: List[int], pdf_class: int) -> int:
        ev = {-1: pdf_class}
        for i, p in enumerate(context):
            ev[i] = p
        n = self.root
        while True:
            t, key, a, b = self.nodes[n]
            if t == 0:
                return a
            if key not in ev:
                raise KeyError(f"event key {key} missing")
            v = ev[key]
            if t == 1:
                n = a if v in self._set_lookup(n) else b
            else:
                ch = self.tables[n]
                if v < 0 or v >= len(ch) or ch[v] < 0:
                    return -1
                n = ch[v]

    def _set_lookup(self, n):
        s = self.sets[n]
        if not isinstance(s, frozenset):
            s = frozenset(int(x) for x in s)
            self.sets[n] = s
        return s

    def flatten(self):
        """Arrays for the C ABI: node[int32 x4], aux offsets + aux pool (SE yes-sets / TE children)."""
        node = np.zeros((len(self.nodes), 4), dtype=np.int32)
        aux_off = np.zeros(len(self.nodes) + 1, dtype=np.int32)
        pool: List[int] = []
        for i, (t, key, a, b) in enumerate(self.nodes):
            node[i] = (t, key, a, b)
            aux_off[i] = len(pool)
            if t == 1:
                pool.extend(sorted(int(x) for x in self.sets[i]))
            elif t == 2:
                pool.extend(self.tables[i])
        aux_off[len(self.nodes)] = len(pool)
        return node, aux_off, np.asarray(pool, dtype=np.int32), self.root


def _read_event_map(r: _Reader, cd: ContextDependency) -> int:
    tok = r.token()
    if tok == "NULL":
        return -1
    idx = len(cd.nodes)
    cd.nodes.append((0, 0, 0, 0))
    if tok == "CE":
        cd.nodes[idx] = (0, 0, r.int32(), 0)
    elif tok == "SE":
        key = r.int32()
        yes = r.int_vector()
        r.expect("{")
        y = _read_event_map(r, cd)
        n = _read_event_map(r, cd)
        r.expect("}")
        cd.nodes[idx] = (1, key, y, n)
        cd.sets[idx] = yes
    elif tok == "TE":
        key = r.int32()
        size = r.int32()
        r.expect("(")
        ch = [_read_event_map(r, cd) for _ in range(size)]
        r.expect(")")
        cd.nodes[idx] = (2, key, 0, 0)
        cd.tables[idx] = ch
    else:
        raise ValueError(f"bad EventMap token {tok!r}")
    return idx


def read_tree(path) -> ContextDependency:
    with open(path, "rb") as f:
        if f.read(2) != b"\0B":
            raise ValueError("tree must be binary")
        r = _Reader(f)
        r.expect("ContextDependency")
        n = r.int32()
        p = r.int32()
        r.expect("ToPdf")
        cd = ContextDependency(n, p)
        cd.root = _read_event_map(r, cd)
        r.expect("EndContextDependency")
    return cd


def monophone_tree(tm_topo: Topology, phone_pdf_classes: Dict[int, List[int]]) -> Tuple[ContextDependency, int]:
    """Build the N=1,P=0 tree ``gmm_init_mono`` would (one pdf per (phone, pdf_class)); returns (tree, num_pdfs)."""
    cd = ContextDependency(1, 0)
    cd.nodes.append((2, 0, 0, 0))
    maxp = int(max(phone_pdf_classes)) + 1
    children = [-1] * maxp
    pdf = 0
    for ph in sorted(phone_pdf_classes):
        classes = phone_pdf_classes[ph]
        tidx = len(cd.nodes)
        cd.nodes.append((2, -1, 0, 0))
        ch = []
        for _c in range(max(classes) + 1):
            ch.append(len(cd.nodes))
            cd.nodes.append((0, 0, pdf, 0))
            pdf += 1
        cd.tables[tidx] = ch
        children[ph] = tidx
    cd.tables[0] = children
    cd.root = 0
    return cd, pdf


def write_tree(path, cd: ContextDependency):
    def w(f, n):
        if n < 0:
            _w_token(f, "NULL")
            return
        t, key, a, b = cd.nodes[n]
        if t == 0:
            _w_token(f, "CE")
            _w_int32(f, a)
        elif t == 1:
            _w_token(f, "SE")
            _w_int32(f, key)
            write_int_vector(f, np.asarray(sorted(int(x) for x in cd.sets[n]), dtype=np.int32))
            _w_token(f, "{")
            w(f, a)
            w(f, b)
            _w_token(f, "}")
        else:
            _w_token(f, "TE")
            _w_int32(f, key)
            f.write(b"\xfc" + struct.pack("<I", len(cd.tables[n])))
            _w_token(f, "(")
            for c in cd.tables[n]:
                w(f, c)
            _w_token(f, ")")

    with open(path, "wb") as f:
        f.write(b"\0B")
        _w_token(f, "ContextDependency")
        _w_int32(f, cd.N)
        _w_int32(f, cd.P)
        _w_token(f, "ToPdf")
        w(f, cd.root)
        _w_token(f, "EndContextDependency")


def read_matrix_file(path) -> np.ndarray:
    with open(path, "rb") as f:
        if f.read(2) != b"\0B":
            raise ValueError("binary matrix expected")
        return _Reader(f).matrix()


def write_matrix_file(path, m: np.ndarray):
    with open(path, "wb") as f:
        f.write(b"\0B")
        write_matrix(f, m)


# --------------------------------------------------------------------------- OpenFst VectorFst<StdArc>
FST_MAGIC = 0x7EB2FDD6
_STATE_HDR = struct.Struct("<fq")
_ARC_DTYPE = np.dtype([("i", "<i4"), ("o", "<i4"), ("w", "<f4"), ("n", "<i4")])


@dataclass
class Fst:
    """Arc-list FST (tropical): arcs[:,0..3] = src, ilabel, olabel, dst; weights f32; finals f32 (inf=non-final)."""
    start: int
    num_states: int
    arc_src: np.ndarray
    arc_ilabel: np.ndarray
    arc_olabel: np.ndarray
    arc_dst: np.ndarray
    arc_weight: np.ndarray
    finals: np.ndarray

    def Start(self) -> int:
        return self.start

    def NumStates(self) -> int:
        return self.num_states


def read_fst(f: BinaryIO) -> Fst:
    magic = struct.unpack("<i", f.read(4))[0]
    if magic & 0xFFFFFFFF != FST_MAGIC:
        raise ValueError("bad OpenFst magic")

    def s():
        n = struct.unpack("<i", f.read(4))[0]
        return f.read(n).decode()

    ftype, atype = s(), s()
    if ftype != "vector" or atype != "standard":
        raise ValueError(f"unsupported fst {ftype}/{atype}")
    _version, flags = struct.unpack("<ii", f.read(8))
    _props, start, nstates, _narcs = struct.unpack("<Qqqq", f.read(32))
    if flags & 3:
        raise ValueError("embedded symbol tables not supported")
    # per state: final weight (f32), arc count (i64), then 16-byte arc records.  Two reads per state; the arc bytes are joined and decoded
    # with ONE frombuffer (a numpy call per state made this parser the slowest stage of the file-based flow)
    finals = np.empty(nstates, dtype=np.float32)
    nas = np.empty(nstates, dtype=np.int64)
    chunks = []
    unpack = _STATE_HDR.unpack
    read = f.read
    for st in range(nstates):
        fw, na = unpack(read(12))
        finals[st] = fw
        nas[st] = na
        if na:
            chunks.append(read(16 * na))
    rec = np.frombuffer(b"".join(chunks), dtype=_ARC_DTYPE)
    src = np.repeat(np.arange(nstates, dtype=np.int32), nas)
    return Fst(int(start), int(nstates), src, rec["i"].copy(), rec["o"].copy(), rec["n"].copy(), rec["w"].copy(), finals)


def read_fst_ark(path) -> List[Tuple[str, "Fst"]]:
    """All (key, Fst) of a binary FST archive (kalpy FstArchive's file).  The archive is read once; each FST's header is parsed here and
    its state / arc body by the C library (mfa_fst_body_scan / _fill): the per-state Python loop of read_fst was the slowest stage of the
    file-based alignment flow (1.1 ms per graph)."""
    import ctypes as C
    from . import _lib as L
    lib = L.lib()
    data = np.fromfile(str(path), dtype=np.uint8)
    raw = data.tobytes()
    base = data.ctypes.data
    n, pos = len(raw), 0
    out: List[Tuple[str, Fst]] = []
    na, nb = C.c_int64(), C.c_int64()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    while pos < n:
        sp = raw.find(b" ", pos)
        if sp < 0:
            break
        key = raw[pos:sp].decode("utf8")
        pos = sp + 1
        if struct.unpack_from("<i", raw, pos)[0] & 0xFFFFFFFF != FST_MAGIC:
            raise ValueError("bad OpenFst magic")
        pos += 4
        types = []
        for _ in range(2):
            ln = struct.unpack_from("<i", raw, pos)[0]
            types.append(raw[pos + 4:pos + 4 + ln].decode())
            pos += 4 + ln
        if types != ["vector", "standard"]:
            raise ValueError(f"unsupported fst {types[0]}/{types[1]}")
        _version, flags = struct.unpack_from("<ii", raw, pos)
        _props, start, nstates, _narcs = struct.unpack_from("<Qqqq", raw, pos + 8)
        pos += 40
        if flags & 3:
            raise ValueError("embedded symbol tables not supported")
        L.check(lib.mfa_fst_body_scan(C.c_void_p(base + pos), C.c_int64(n - pos), C.c_int64(nstates), C.byref(na), C.byref(nb)))
        A = int(na.value)
        finals = np.empty(nstates, np.float32)
        src, dst, il, ol = (np.empty(A, np.int32) for _ in range(4))
        w = np.empty(A, np.float32)
        L.check(lib.mfa_fst_body_fill(C.c_void_p(base + pos), C.c_int64(nstates), vp(finals), vp(src), vp(dst), vp(il), vp(ol), vp(w)))
        pos += int(nb.value)
        out.append((key, Fst(int(start), int(nstates), src, il, ol, dst, w, finals)))
    return out


def write_fst(f: BinaryIO, fst: Fst):
    f.write(struct.pack("<I", FST_MAGIC))
    for s in ("vector", "standard"):
        f.write(struct.pack("<i", len(s)) + s.encode())
    f.write(struct.pack("<ii", 2, 0))
    f.write(struct.pack("<Qqqq", 0x0000000000000003, fst.start, fst.num_states, 0))
    order = np.argsort(fst.arc_src, kind="stable")
    src = fst.arc_src[order]
    rec = np.zeros(order.shape[0], dtype=_ARC_DTYPE)
    rec["i"], rec["o"], rec["w"], rec["n"] = fst.arc_ilabel[order], fst.arc_olabel[order], fst.arc_weight[order], fst.arc_dst[order]
    S = fst.num_states
    bounds = np.searchsorted(src, np.arange(S + 1))
    # one byte image: 12-byte state headers interleaved with the arc records (no per-state writes)
    na = np.diff(bounds).astype(np.int64)
    body = np.zeros(12 * S + 16 * int(order.shape[0]), dtype=np.uint8)
    hdr_off = 12 * np.arange(S, dtype=np.int64) + 16 * bounds[:-1].astype(np.int64)
    hdr = np.zeros(S, dtype=[("f", "<f4"), ("n", "<i8")])
    hdr["f"], hdr["n"] = np.asarray(fst.finals, np.float32), na
    hb = hdr.view(np.uint8).reshape(S, 12)
    body[(hdr_off[:, None] + np.arange(12)).ravel()] = hb.ravel()
    if order.shape[0]:
        arc_off = np.repeat(hdr_off + 12 - 16 * bounds[:-1].astype(np.int64), na) + 16 * np.arange(order.shape[0], dtype=np.int64)
        body[(arc_off[:, None] + np.arange(16)).ravel()] = rec.view(np.uint8).reshape(-1, 16).ravel()
    f.write(body.tobytes())


# --------------------------------------------------------------------------- ark / scp tables
def _read_key(f: BinaryIO) -> Optional[str]:
    out = bytearray()
    while True:
        c = f.read(1)
        if not c:
            return None if not out else out.decode("utf8")
        if c == b" ":
            return out.decode("utf8")
        out += c


def _read_object(f: BinaryIO, kind: str):
    if kind == "fst":
        return read_fst(f)
    if f.read(2) != b"\0B":
        raise ValueError("text-mode archives not supported")
    r = _Reader(f)
    if kind == "int_vector":
        return r.int_vector()
    if kind == "vector":
        return r.vector()
    if kind == "matrix":
        return r.matrix()
    raise ValueError(kind)


def read_ark(path, kind: str) -> Iterator[Tuple[str, object]]:
    """Iterate (key, object) over a Kaldi archive. kind in int_vector|vector|matrix|fst."""
    with open(path, "rb") as f:
        while True:
            k = _read_key(f)
            if k is None:
                return
            yield k, _read_object(f, kind)


def read_scp(path) -> List[Tuple[str, str, int]]:
    out = []
    with open(path, "r", encoding="utf8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            key, rest = line.split(None, 1)
            p, _, off = rest.rpartition(":")
            out.append((key, p, int(off)))
    return out


def read_scp_object(path: str, offset: int, kind: str):
    with open(path, "rb") as f:
        f.seek(offset)
        return _read_object(f, kind)


class ArkWriter:
    """``ark[,scp]`` writer (kalpy ``*Writer`` / ``generate_write_specifier`` equivalent)."""

    def __init__(self, ark_path, scp_path=None):
        self.ark_path = str(ark_path)
        self.f = open(ark_path, "wb")
        self.scp = open(scp_path, "w", encoding="utf8") if scp_path else None

    def _key(self, key: str):
        self.f.write(key.encode("utf8") + b" ")
        if self.scp:
            self.scp.write(f"{key} {self.ark_path}:{self.f.tell()}\n")

    def write_int_vector(self, key, v):
        self._key(key)
        self.f.write(b"\0B")
        write_int_vector(self.f, v)

    def write_vector(self, key, v):
        self._key(key)
        self.f.write(b"\0B")
        write_vector(self.f, v)

    def write_matrix(self, key, m, compress=False):
        self._key(key)
        self.f.write(b"\0B")
        if compress:
            self.f.write(compress_matrix(m))
        else:
            write_matrix(self.f, m)

    def write_fst(self, key, fst: Fst):
        self._key(key)
        write_fst(self.f, fst)

    def close(self):
        self.f.close()
        if self.scp:
            self.scp.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_wav_int16(path, channel: int = 0) -> Tuple[np.ndarray, int]:
    """Minimal RIFF/WAVE PCM reader (16-bit int; 24/32-bit int and float32 are rescaled to the int16 range,
    as kalpy's Segment does after librosa.load)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos = 12
    fmt = None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        sz = struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + sz]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            tag, nch, sr, _, _, bits = fmt
            if tag == 1 and bits == 16:
                x = np.frombuffer(body[: len(body) // 2 * 2], dtype="<i2").astype(np.float32)
            elif tag == 1 and bits == 24:
                b = np.frombuffer(body[: len(body) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
                v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
                v = np.where(v >= 1 << 23, v - (1 << 24), v)
                x = v.astype(np.float32) / 256.0
            elif tag == 1 and bits == 32:
                x = np.frombuffer(body[: len(body) // 4 * 4], dtype="<i4").astype(np.float32) / 65536.0
            elif tag == 3 and bits == 32:
                x = np.frombuffer(body[: len(body) // 4 * 4], dtype="<f4") * 32768.0
            else:
                raise ValueError(f"unsupported wav format tag={tag} bits={bits}")
            x = x.reshape(-1, nch)[:, channel]
            return np.round(x).clip(-32768, 32767).astype(np.int16), sr
        pos += 8 + sz + (sz & 1)
    raise ValueError("no data chunk")


def write_wav_int16(path, pcm: np.ndarray, sample_rate: int = 16000):
    """Mono 16-bit RIFF/WAVE writer (tests and synthetic corpora)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + pcm.nbytes) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, sample_rate, 2 * sample_rate, 2, 16))
        f.write(b"data" + struct.pack("<I", pcm.nbytes) + pcm.tobytes())
