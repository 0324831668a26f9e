"""In-tree build of libflagstats_cuda.so for sm_100a.

    python -m libflagstats_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so lands next to this file
(libflagstats_b200/libflagstats_cuda.so), is git-ignored and travels to the
GPU box inside the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libflagstats_cuda.so")

SOURCES = ["flagstat_capi.cu"]
HEADERS = ["flagstat_kernels.cuh", "flagstat_kernel_tma.cuh", "flagstat_kernel_group.cuh", "flagstat_kernel_dyn.cuh", "lz4_block.cuh", "lz4_block_group.cuh", "lz4_block_cta.cuh", "zstd_frame.cuh", "zstd_block.cuh", "ingest_text.cuh", "flagstat_blockfile.inl", "bitcounter.cuh", "synth.cuh",
           os.path.join("..", "..", "include", "flagstats_cuda.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-shared", "-cudart", "shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


VARIANTS_SO = os.path.join(os.path.dirname(HERE), "tools", "bin", "libflagstats_cuda_variants.so")


def build(force: bool = False, verbose: bool = False, all_variants: bool = False) -> str:
    """all_variants: the A/B build with every measured-and-superseded kernel variant compiled in
    (-DFSB_ALL_VARIANTS), written to tools/bin/ -- never the product library."""
    out = VARIANTS_SO if all_variants else SO
    if not all_variants and not force and not stale():
        return SO
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + (["-DFSB_ALL_VARIANTS"] if all_variants else [])
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv,
                all_variants="--all-variants" in sys.argv))
