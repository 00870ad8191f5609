"""Builds adapt_b200/lib/libadapt_b200.so with nvcc for sm_100a (in-tree, so the .so travels to the GPU box).

    python -m adapt_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB = os.path.join(LIB_DIR, "libadapt_b200.so")
SOURCES = ["adapt_abi.cu", "bvh_device.cu", "bvh_build.cpp"]
HEADERS = ["pt_common.cuh", "pt_shade.cuh", "pt_trace.cuh", "pt_path.cuh", "pt_volume.cuh", "pt_kernels.cuh", "scene_pack.h", "bvh_build.h", "bvh_device.h", "bvh_lbvh.h", os.path.join("..", "..", "include", "adapt_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(f) > t for f in files if os.path.exists(f))


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None) -> str:
    """out: alternative output path (tuning variants, e.g. -DLOGIC_MIN_BLOCKS=3; pick one at run time with ADAPT_B200_LIB)."""
    if out is None and not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    out = out or LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-ccbin", "g++", "-Xcompiler", "-fPIC,-fopenmp,-O3", "--shared", "-o", out]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += list(extra_flags)
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-lgomp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
