"""The headless C++17 host mirror (voxelpathtracer_b200/host/VoxelRT.h + headless_main.cpp) builds against the C ABI; without a
GPU it fails loudly, and on a GPU it produces exactly the planes the Python driver produces for the same frames."""
import os
import subprocess

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera, world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "voxelpathtracer_b200")
EXE = os.path.join(PKG, "host", "vxpt_headless")


def build_headless():
    abi.load()  # libvxpt.so must exist
    cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cc, "-std=c++17", "-O2", "-Wall", "-Werror", "-o", EXE, os.path.join(PKG, "host", "headless_main.cpp"), "-L" + PKG, "-lvxpt",
                    "-Wl,-rpath," + PKG], check=True)
    return EXE


def fnv1a(a):
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_host_mirror_compiles_against_the_abi():
    exe = build_headless()
    assert os.path.exists(exe)
    src = open(os.path.join(PKG, "host", "VoxelRT.h")).read()
    for name in ("class World", "GetBlock", "SetBlock", "Buffer", "InitializeDistanceGenerator", "GenerateDistanceField", "GenerateWorld",
                 "SaveWorld", "LoadWorld", "BlockDataSSBO", "BlueNoiseDataSSBO", "FPSCamera", "GetTAAJitter"):
        assert name in src, name   # the reference's names for this path


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_headless_fails_loudly_without_a_gpu():
    exe = build_headless()
    p = subprocess.run([exe, "64", "36"], capture_output=True, text=True)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("use_plains", [False, True])
def test_headless_cpp_equals_python_driver(use_plains, plains_columns):
    exe = build_headless()
    W, H = 160, 90
    args = [exe, str(W), str(H)] + ([os.path.join(ROOT, "tests", "golden", "plains_columns.u8")] if use_plains else [])
    out = subprocess.run(args, capture_output=True, text=True, check=True).stdout.split("\n")
    got = {}
    for line in out:
        f = line.split()
        if not f:
            continue
        if f[0] in ("df", "df_after_edit"):
            got[f[0]] = int(f[1], 16)
        elif f[0] == "frame":
            got[("frame", int(f[1]))] = {f[i]: int(f[i + 1], 16) for i in range(2, len(f), 2)}
        elif f[0] == "render_frame":
            got["render_frame"] = {f[i]: int(f[i + 1], 16) for i in range(1, len(f), 2)}
    w = world.generate_plains(plains_columns) if use_plains else world.generate_superflat()
    r = vx.Renderer(0)
    try:
        r.upload_world(w)
        r.build_distance_field()
        assert fnv1a(r.download_distance_field()) == got["df"]
        cam = camera.FpsCamera(yaw_deg=90.0, pitch_deg=-20.0, aspect=W / H).vx_camera(W, H)
        sun = np.array([-0.66896474, 0.46841538, 0.57735026], np.float32)
        for frame in range(3):
            g = r.trace_primary(cam, vx.primary_params(475 if frame == 0 else 350, camera.taa_jitter(frame)), r.alloc_gbuffer(W, H))
            s = r.trace_shadow(cam, g, vx.shadow_params(sun, frame=frame, soft=False), r.alloc_shadow(W, H))
            ref = got[("frame", frame)]
            assert fnv1a(g["t"]) == ref["t"] and fnv1a(g["normal_id"]) == ref["normal"] and fnv1a(g["block_id"]) == ref["block"]
            assert fnv1a(s["shadow"]) == ref["shadow"]
        assert got["render_frame"] == got[("frame", 2)]   # vxpt_render_frame from C++ == the separate calls
        r.set_block(192, 70, 200, world.STONE)
        r.build_distance_field()
        assert fnv1a(r.download_distance_field()) == got["df_after_edit"]
    finally:
        r.close()
