"""In-tree nvcc build of libpgfuse.so (sm_100a only).  `python -m eeg_multimodal_b200.build`.

The shared library is a plain C-ABI object (include/pgfuse.h): it links against the CUDA
runtime only -- no torch, no pybind -- so any host language can bind it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpgfuse.so")
SOURCES = ["capi.cu", "perturb_gate.cu", "linear_simt.cu", "linear_stream.cu", "linear_wide.cu", "gemm_tc.cu", "fp32x3.cu", "ce_cls.cu", "adam.cu", "prigumbel.cu", "sweep_step.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(out: str, deps: list[str]) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "pgfuse.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            log = open(o + ".log", "w")
            procs.append((src, o, log, subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", s, "-o", o], stdout=log, stderr=subprocess.STDOUT)))
    failed = []
    for src, o, log, p in procs:
        rc = p.wait()
        log.close()
        text = open(o + ".log").read()
        if verbose or rc != 0:
            sys.stderr.write(f"--- nvcc {src}\n{text}\n")
        if rc != 0:
            failed.append(src)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
