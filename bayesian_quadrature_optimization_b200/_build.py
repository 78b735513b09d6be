"""Builds libbqo_sm100.so (sm_100a only) in-tree with nvcc.

The shared library is the product's only compute path; there is no CPU fallback.  nvcc
cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "_lib"
LIB_PATH = LIB_DIR / "libbqo_sm100.so"
STAMP = LIB_DIR / "libbqo_sm100.stamp"

SOURCES = [
    "bqo_assemble.cu",
    "bqo_chol.cu",
    "bqo_loglik_grad.cu",
    "bqo_stage_c.cu",
    "bqo_posterior.cu",
    "bqo_hessian.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "179",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libbqo_sm100.so cannot be built")
    return exe


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh"))
    files.append(PKG_DIR.parent / "include" / "bqo_sm100.h")
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the C-ABI shared library."""
    LIB_DIR.mkdir(exist_ok=True)
    digest = _source_hash()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = LIB_DIR / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode()))
        if verbose and out:
            print(out.decode())
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout.decode())
    STAMP.write_text(digest)
    return LIB_PATH


def build_variant(tag: str, defines: dict) -> Path:
    """Tuning aid: build libbqo_sm100_<tag>.so with extra -D macros (e.g. IM_WARPS, EVAL_NV).
    Select it at run time with the environment variable BQO_LIB_PATH."""
    LIB_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    out = LIB_DIR / ("libbqo_sm100_%s.so" % tag)
    objs = []
    # macros that only bqo_stage_c.cu reads recompile that file alone; anything else, everything
    stage_c_only = all(k in ("IM_WARPS", "EVAL_NV", "BQO_EXP_BITS", "BQO_POLY_UR", "BQO_HALF_INT")
                       for k in defines)
    flags = ["-D%s=%s" % kv for kv in defines.items()]
    procs = []
    for src in SOURCES:
        obj = LIB_DIR / (src[:-3] + ".o")
        if src == "bqo_stage_c.cu" or not stage_c_only:
            obj = LIB_DIR / ("%s_%s.o" % (src[:-3], tag))
            cmd = [nvcc, *NVCC_FLAGS, *flags, "-c", str(CSRC / src), "-o", str(obj)]
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        objs.append(str(obj))
    for pr in procs:
        out_, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(out_.decode())
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError(r.stdout.decode())
    return out


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        # python _build.py --variant w20n2 IM_WARPS=20 EVAL_NV=2
        build(force=False)
        print(build_variant(sys.argv[2], dict(a.split("=") for a in sys.argv[3:])))
    else:
        print(build(force=True, verbose=True))
