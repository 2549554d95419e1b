"""Builder of tests/host_shadow/libvxpt_hostemu.so (vxpt_hostemu.cpp): the C ABI compiled by g++ against a miniature CUDA runtime.
TEST INFRASTRUCTURE ONLY — used by `pytest --host-emulation` (tests/conftest.py) to run the host-plane GPU tests on a machine without a
GPU; never imported by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, "libvxpt_hostemu.so")
CSRC = os.path.join(ROOT, "voxelpathtracer_b200", "csrc")
CUDA_INCLUDE = "/usr/local/cuda/include"


def build(force=False):
    srcs = [os.path.join(HERE, "vxpt_hostemu.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "vxpt.h"), os.path.join(HERE, "warp_emu.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs):
        return LIB_PATH
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fopenmp", "-shared", "-fvisibility=hidden",
                           "-Wno-attributes", "-Wno-unknown-pragmas", "-I" + CUDA_INCLUDE, "-x", "c++", srcs[0], "-o", LIB_PATH], cwd=HERE)
    return LIB_PATH
