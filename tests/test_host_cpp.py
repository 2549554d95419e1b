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
                 "SaveWorld", "LoadWorld", "BlockDataSSBO", "BlueNoiseDataSSBO", "FPSCamera", "GetTAAJitter", "Raycast", "RaycastDetect",
                 "GetViewProjection"):
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
    args = [exe, str(W), str(H)] + ([os.path.join(ROOT, "voxelpathtracer_b200", "data", "plains_columns.u8")] if use_plains else [])
    # the C++ host also shards the last frame over three handles through vxpt_mg_* (three slabs on device 0: runs on a single-GPU box)
    out = subprocess.run(args, capture_output=True, text=True, check=True, env=dict(os.environ, VXPT_HEADLESS_DEVICES="0,0,0")).stdout.split("\n")
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
        elif f[0] == "mg_render_frame":
            got["mg_devices"] = int(f[2])
            got["mg_render_frame"] = {f[i]: int(f[i + 1], 16) for i in range(3, len(f), 2)}
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
        assert got["mg_devices"] == 3 and got["mg_render_frame"] == got[("frame", 2)]   # ... == the frame sharded over three handles by vxpt_mg_*
        r.set_block(192, 70, 200, world.STONE)
        r.build_distance_field()
        assert fnv1a(r.download_distance_field()) == got["df_after_edit"]
    finally:
        r.close()


def test_view_projection_of_the_two_host_mirrors():
    """camera.FpsCamera.view_projection_f32 == the textbook lookAt / perspective to float precision, and inverts the inv_view / inv_proj the
    trace passes receive (the C++ mirror's GetViewProjection is the same code line for line; the GPU test below compares their effects)."""
    fc = camera.FpsCamera(position=(100.25, 61.5, 99.75), yaw_deg=37.0, pitch_deg=-12.0, aspect=160 / 90)
    view, proj = fc.view_projection_f32()
    assert np.allclose(view.reshape(4, 4).T, fc.view(), atol=2e-5) and np.allclose(proj.reshape(4, 4).T, fc.projection(), atol=1e-6)
    iv, ip = fc.inverse_matrices()
    assert np.allclose(view.reshape(4, 4).T.astype(np.float64) @ iv.astype(np.float64), np.eye(4), atol=1e-5)
    assert np.allclose(proj.reshape(4, 4).T.astype(np.float64) @ ip.astype(np.float64), np.eye(4), atol=1e-5)


def test_camera_matrices_of_the_two_host_mirrors_are_bit_identical(tmp_path):
    abi.load()
    cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "camera")
    subprocess.run([cc, "-std=c++17", "-O2", "-Wall", "-Werror", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "camera_main.cpp"),
                    "-L" + PKG, "-lvxpt", "-Wl,-rpath," + PKG], check=True)
    rng = np.random.RandomState(4)
    poses = [(192.0, 75.0, 192.0, 90.0, -20.0, 16.0 / 9.0)] + [tuple(np.float32(v) for v in (rng.uniform(1, 380), rng.uniform(1, 120), rng.uniform(1, 380),
                                                                                          rng.uniform(-180, 180), rng.uniform(-85, 85))) + (float(rng.choice([16 / 9, 4 / 3, 160 / 90, 1.0])),)
                                                               for _ in range(40)]
    text = "\n".join("%.9g %.9g %.9g %.9g %.9g %.17g" % p for p in poses) + "\n"
    lines = [ln for ln in subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split("\n") if ln.strip()]
    assert len(lines) == len(poses)
    for p, ln in zip(poses, lines):
        got = np.array([int(h, 16) for h in ln.split()], dtype=np.uint32).view(np.float32).reshape(4, 16)
        fc = camera.FpsCamera(position=p[:3], yaw_deg=float(p[3]), pitch_deg=float(p[4]), aspect=float(p[5]))
        view, proj = fc.view_projection_f32()
        cam = fc.vx_camera(64, 36)
        assert np.array_equal(got[0], view) and np.array_equal(got[1], proj), p
        assert np.array_equal(got[2], np.array(list(cam.inv_view), np.float32)) and np.array_equal(got[3], np.array(list(cam.inv_proj), np.float32)), p
