"""In-tree build of the CUDA library (libvxpt.so) for sm_100a.  nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["csrc/api.cu", "csrc/df_build.cu", "csrc/trace.cu", "csrc/trace_gi.cu", "csrc/trace_reflection.cu", "csrc/df_consumers.cu", "csrc/gbuffer.cu", "csrc/denoise.cu", "csrc/l2_probe.cu", "csrc/mg.cu"]
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


def build(force=False, verbose=False, defines=(), out=None):
    """Build libvxpt.so.  defines / out: development only — an experiment variant of the library (-DNAME=VALUE ...) under another file
    name, which the measurement tools under tools/ can be pointed at (their VXPT_LIB variable); the product only ever loads libvxpt.so."""
    lib = out or LIB
    if not force and not defines and not needs_build():
        return lib
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + SOURCES
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building " + os.path.basename(lib))
    if verbose:
        sys.stderr.write(proc.stderr)
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=os.path.join(HERE, outs[0]) if outs else None))
