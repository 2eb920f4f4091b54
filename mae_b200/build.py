"""Builds libmae_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m mae_b200.build            # or __graft_entry__.build()

The library is linked against the static CUDA runtime and uses no torch headers: the boundary is
plain C (include/mae_b200.h).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmae_b200.so")
SOURCES = ["api.cu", "reparam_kl.cu", "mpd.cu", "mpd_fused.cu", "mpd_rows_fused.cu", "bernoulli.cu", "dmol.cu", "dmol_head.cu", "logprob_iw.cu", "assemble.cu", "p2p_gather.cu", "microbench.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libmae_b200.so")
    return exe


def _extra_flags():
    extra = ["-DMAE_DEBUG_TIMING"] if os.environ.get("MAE_DEBUG_TIMING") else []   # phase timestamps, profiling builds only
    if os.environ.get("MAE_DMOL_G5"):                                              # tuning: pixel groups per CTA (SPLIT=5)
        extra.append("-DMAE_DMOL_G5=%d" % int(os.environ["MAE_DMOL_G5"]))
    if os.environ.get("MAE_DMOL_BWD_MINB"):                                        # tuning: min CTAs/SM of the DMOL backward
        extra.append("-DMAE_DMOL_BWD_MINB=%d" % int(os.environ["MAE_DMOL_BWD_MINB"]))
    return extra


def _flags_file() -> str:
    return os.path.join(PKG, "build", "flags.txt")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    try:
        with open(_flags_file()) as f:
            if f.read() != " ".join(_extra_flags()):
                return True
    except OSError:
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "mae_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into mae_b200/libmae_b200.so; returns the path."""
    if not force and not _stale():
        return LIB
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    extra = _extra_flags()
    objs = []
    logs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, cmd, p in procs:
        out, _ = p.communicate()
        logs.append("### %s\n%s" % (src, out))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, " ".join(cmd), out))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    with open(_flags_file(), "w") as f:
        f.write(" ".join(extra))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
