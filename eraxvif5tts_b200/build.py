"""Build libf5b200.so (sm_100a only) in-tree with nvcc.   python -m eraxvif5tts_b200.build [--force]

One object per translation unit (compiled in parallel, rebuilt when the source or a header is newer), linked into
eraxvif5tts_b200/lib/libf5b200.so.  nvcc cross-compiles without a GPU; the built .so travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
# F5B_BUILD_TAG=<tag> (with F5B_NVCC_EXTRA=-D...): an instrumented / experimental build next to the product library
# (lib/libf5b200_<tag>.so, loaded with F5B_LIB=...), never instead of it
_TAG = os.environ.get("F5B_BUILD_TAG", "")
OBJDIR = os.path.join(PKG, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(LIBDIR, "libf5b200" + ("_" + _TAG if _TAG else "") + ".so")
UNITS = ["host", "gemm", "gemm_grad", "convpos", "attention", "attention_tf32", "attention_bwd", "elementwise", "spectral", "optim", "train_kernels", "train", "align", "models"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libf5b200.so cannot be built (set NVCC=/path/to/nvcc)")


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "f5b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    extra = os.environ.get("F5B_NVCC_EXTRA", "").split()  # e.g. -DATT_TRACE for the instrumented attention build
    global UNITS
    if "-DF5B_WITH_ATTN_FA" in extra and "attention_fa" not in UNITS:
        # the experimental persistent attention forward (slower than attention.cu; kept for its traces and notes,
        # profiles/r02_attention_fa_notes.md) is only part of tagged experiment builds, never of the product library
        UNITS = UNITS + ["attention_fa"]
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hm = _headers_mtime()
    todo = []
    for u in UNITS:
        src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(OBJDIR, u + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            todo.append((src, obj))

    def compile_one(so):
        src, obj = so
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            list(ex.map(compile_one, todo))
    objs = [os.path.join(OBJDIR, u + ".o") for u in UNITS]
    if todo or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
