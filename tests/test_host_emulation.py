"""The C ABI on the CPU: tests/host_shadow/vxpt_hostemu.cpp compiles voxelpathtracer_b200/csrc/api.cu (argument checks, plane staging, slab
arithmetic, call order) with g++ against a miniature CUDA runtime and runs the per-pixel kernels' own source thread after thread.  This
test runs the `-m gpu` tests of the newer exports that use HOST planes against that library, in a subprocess (`pytest --host-emulation`),
so that a mistake in an export or in its GPU test shows up in the CPU suite instead of at the end of a round.  It proves nothing about
the GPU itself and is test infrastructure only: the product library is nvcc's build and has no host path."""
import os
import subprocess
import sys

import pytest

from host_shadow import kernels_on_host as koh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not koh.available(), reason="CUDA toolkit headers not present")
def test_host_plane_gpu_tests_pass_against_the_emulated_abi():
    files = ["tests/test_material_pass.py", "tests/test_per_frame_edits.py", "tests/test_svgf_denoise.py", "tests/test_alpha_traversal.py",
             "tests/test_z_material_extras.py", "tests/test_zz_material_quad_shuffle.py", "tests/test_mg_frame.py"]
    # device-resident planes need torch's CUDA allocator, the headless C++ binary links the real library: left to the GPU box
    select = "not True and not device and not headless and not whole_chain"
    p = subprocess.run([sys.executable, "-m", "pytest", *files, "-m", "gpu", "--host-emulation", "-q", "-x", "-k", select, "-p", "no:cacheprovider"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = "\n".join(p.stdout.split("\n")[-15:])
    assert p.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
