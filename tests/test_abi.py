"""The C-ABI library loads, exports every symbol include/vxpt.h declares, its structs have the layout the ctypes
binding assumes, and it fails loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from voxelpathtracer_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vxpt.h")


def declared_exports():
    src = open(HEADER).read()
    return re.findall(r"VXPT_API\s+[\w\s\*]+?\b(vxpt_\w+)\s*\(", src)


def test_header_declares_the_expected_surface():
    names = declared_exports()
    assert len(names) == len(set(names)) >= 27
    for required in ("vxpt_create", "vxpt_upload_world", "vxpt_set_block", "vxpt_build_distance_field", "vxpt_download_distance_field",
                     "vxpt_set_materials", "vxpt_set_blue_noise", "vxpt_trace_primary", "vxpt_trace_shadow", "vxpt_trace_diffuse",
                     "vxpt_trace_reflection", "vxpt_sync", "vxpt_last_error", "vxpt_get_stats"):
        assert required in names


def test_binding_covers_exactly_the_header():
    assert sorted(declared_exports()) == sorted(abi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = abi.load()  # raises if libvxpt.so is missing or a symbol cannot be bound
    for name in declared_exports():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(\w+)", out))
    assert set(declared_exports()) <= exported
    # nothing but the ABI leaks out of the shared object
    assert {n for n in exported if not n.startswith("vxpt_")} <= {"_init", "_fini"}
    assert lib.vxpt_version().decode().startswith("vxpt ")


def test_header_every_export_cites_a_reference_site():
    src = open(HEADER).read()
    for path in ("Core/World.cpp", "Core/World.h", "Core/Pipeline.cpp", "Core/BlockDataSSBO.cpp", "Core/BlueNoiseDataSSBO.cpp",
                 "InitialRayTraceFrag.glsl", "ShadowRayTraceFrag.glsl", "DiffuseRayTraceFrag.glsl", "ReflectionTraceFrag.glsl"):
        assert path in src, path


def test_integration_guide_indexes_every_export():
    """INTEGRATION.md section 7 names the reference site each export stands in for: the table has to cover the header exactly."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    index = text.split("## 7. Index", 1)[1]
    assert sorted(set(re.findall(r"`(vxpt_\w+)`", index))) == sorted(declared_exports())


def test_struct_layouts_match_the_c_compiler():
    structs = ["VxCamera", "VxPrimaryParams", "VxGBuffer", "VxShadowParams", "VxShadowOut", "VxDiffuseParams", "VxDiffuseOut",
               "VxReflectionParams", "VxReflectionIn", "VxReflectionOut", "VxStats", "VxFrameParams", "VxFrameOut", "VxMaterialParams", "VxMaterialOut",
               "VxSvgfInitialIn", "VxSvgfInitialOut", "VxSvgfTemporalIn", "VxSvgfTemporalParams", "VxSvgfTemporalOut", "VxSvgfVarianceIn",
               "VxSvgfVarianceParams", "VxSvgfVarianceOut", "VxSvgfSpatialIn", "VxSvgfSpatialParams", "VxSvgfSpatialOut", "VxSvgfFrameParams",
               "VxShadowTemporalIn", "VxShadowTemporalParams", "VxShadowTemporalOut", "VxShadowFilterIn", "VxShadowFilterParams", "VxShadowFrameParams"]
    prog = '#include <stdio.h>\n#include "vxpt.h"\nint main(void){' + "".join(
        f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + \
        'printf("off_row_begin %zu\\n", __builtin_offsetof(VxCamera,row_begin));' \
        'printf("off_sun_dir %zu\\n", __builtin_offsetof(VxDiffuseParams,sun_dir));' \
        'printf("off_material %zu\\n", __builtin_offsetof(VxFrameParams,material));' \
        'printf("off_time %zu\\n", __builtin_offsetof(VxSvgfFrameParams,time));' \
        'printf("off_prev_sh %zu\\n", __builtin_offsetof(VxSvgfTemporalIn,prev_sh));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(cfile, "w").write(prog)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-std=c99", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)  # header is valid C
        out = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for s in structs:
        assert int(out[s]) == C.sizeof(getattr(abi, s)), s
    assert int(out["off_row_begin"]) == abi.VxCamera.row_begin.offset
    assert int(out["off_sun_dir"]) == abi.VxDiffuseParams.sun_dir.offset
    assert int(out["off_material"]) == abi.VxFrameParams.material.offset
    assert int(out["off_time"]) == abi.VxSvgfFrameParams.time.offset
    assert int(out["off_prev_sh"]) == abi.VxSvgfTemporalIn.prev_sh.offset


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_create_fails_loudly_without_a_gpu():
    lib = abi.load()
    h = C.c_void_p()
    rc = lib.vxpt_create(0, C.byref(h))
    assert rc == abi.E_CUDA and not h.value
    assert b"no CPU fallback" in lib.vxpt_last_error()
    import voxelpathtracer_b200 as vx
    with pytest.raises(abi.VxptError):
        vx.Renderer(0)


def test_null_arguments_are_rejected_not_crashed():
    lib = abi.load()
    assert lib.vxpt_create(0, None) == abi.E_INVALID
    assert lib.vxpt_sync(None) == abi.E_INVALID
    assert lib.vxpt_upload_world(None, None) == abi.E_INVALID
    assert lib.vxpt_build_distance_field(None) == abi.E_INVALID
    assert lib.vxpt_get_stats(None, None) == abi.E_INVALID
    assert lib.vxpt_destroy(None) == abi.OK
    assert lib.vxpt_last_error() is not None


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under voxelpathtracer_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "voxelpathtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "vxo_" not in text and "libvxo" not in text and "from oracle" not in text and "import oracle" not in text, f
                # ... nor the g++ builds of its sources that tests/host_shadow keeps for the CPU suite (kernels on the host, the emulated ABI)
                assert "hostemu" not in text and "kernels_on_host" not in text and "import host_shadow" not in text and "from host_shadow" not in text, f
    out = subprocess.run(["ldd", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "vxo" not in out and "hostemu" not in out
    # the binding has exactly one library path, next to the package, and no environment override
    src = open(os.path.join(pkg, "abi.py")).read()
    assert src.count("LIB_PATH =") == 1 and "environ" not in src
