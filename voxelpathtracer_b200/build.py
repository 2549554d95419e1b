"""In-tree build of the CUDA library (libvxpt.so) for sm_100a.  nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/api.cu", "csrc/df_build.cu", "csrc/trace.cu", "csrc/trace_gi.cu", "csrc/trace_reflection.cu", "csrc/df_consumers.cu", "csrc/gbuffer.cu", "csrc/denoise.cu", "csrc/l2_probe.cu"]
HEADERS = ["csrc/vxpt_internal.h", "csrc/trace_device.cuh", "csrc/gi_device.cuh", "../include/vxpt.h"]
LIB = os.path.join(HERE, "libvxpt.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exact parity with the reference arithmetic: no FMA contraction, IEEE division and square root
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libvxpt.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
