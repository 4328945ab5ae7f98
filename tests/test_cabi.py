"""The drop-in boundary: libmfa_b200.so loads, exports every symbol include/mfa_b200.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

from mfa_b200 import _lib as L, engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "mfa_b200.h")).read()
    declared = re.findall(r"MFA_API\s+[\w\s\*]+?\b(mfa_\w+)\s*\(", hdr)
    assert len(declared) >= 30 and len(set(declared)) == len(declared)
    assert sorted(declared) == sorted(L.SYMBOLS)
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mfa_abi_version() == 2


def test_struct_layouts_match_header_order():
    # field counts of the ctypes mirrors == the header's structs (a cheap guard against silent ABI drift)
    hdr = open(os.path.join(ROOT, "include", "mfa_b200.h")).read()
    for cname, cls in (("mfa_mfcc_opts", L.MfccOpts), ("mfa_feat_opts", L.FeatOpts), ("mfa_model_desc", L.ModelDesc),
                       ("mfa_hmm_desc", L.HmmDesc), ("mfa_lexicon_desc", L.LexiconDesc), ("mfa_align_opts", L.AlignOpts),
                       ("mfa_trans_desc", L.TransDesc), ("mfa_mle_opts", L.MleOpts), ("mfa_mle_result", L.MleResult)):
        body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\}\s*" + cname + ";", hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        n = sum(len(decl.split(",")) for decl in body.split(";") if decl.strip())
        assert n == len(cls._fields_), (cname, n, len(cls._fields_))


def test_num_frames_host_function():
    o = E.mfcc_opts()
    assert E.num_frames(o, 399) == 0 and E.num_frames(o, 400) == 1 and E.num_frames(o, 427572) == 2670
    assert E.num_frames(E.mfcc_opts(snip_edges=False), 427572) == 2672
    with pytest.raises(L.MfaError):
        E.mfcc_opts(dither=1.0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.MfaError, match="no usable CUDA device"):
        E.Engine(0)
