"""In-tree build of libmfa_b200.so (CUDA kernels + C ABI + host graph compiler) for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmfa_b200.so")
SOURCES = ["engine.cu", "mfcc.cu", "feats.cu", "gmm.cu", "gmm_tc.cu", "viterbi.cu", "viterbi_band.cu", "accstats.cu", "fmllr.cu", "mstep.cu", "pipeline.cu", "graph.cc"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mfa_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None, objdir: str = None) -> str:
    """`extra_flags` / `out` / `objdir`: development variants (e.g. -DMFA_TC_EXP=3 into /tmp) that never replace the in-tree library."""
    if out is None and not force and not _stale():
        return LIB
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC, *FLAGS, *extra_flags, "-c", os.path.join(CSRC, s), "-o", o]
        if s.endswith(".cc"):
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{log}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", out or LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
