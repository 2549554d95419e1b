import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--host-emulation", action="store_true", default=False,
                     help="development aid for machines without a GPU: bind the ctypes layer to tests/host_shadow/libvxpt_hostemu.so (the C ABI "
                          "compiled by g++ against a miniature CUDA runtime, kernels run thread after thread) and run the gpu-marked tests "
                          "that use host planes.  Says nothing about the GPU; never used by the driver's runs.")


_EMULATED = False


def pytest_configure(config):
    global _EMULATED
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--host-emulation"):
        from host_shadow import hostemu
        from voxelpathtracer_b200 import abi
        abi.LIB_PATH = hostemu.build()
        abi._lib = None
        _EMULATED = True


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu() or _EMULATED:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---------------------------------------------------------------------------------------------- shared fixtures
@pytest.fixture(scope="session")
def plains_columns():
    from voxelpathtracer_b200 import assets
    return assets.load_plains_columns()


@pytest.fixture(scope="session")
def worlds(plains_columns):
    """name -> World, built lazily and cached for the session."""
    from voxelpathtracer_b200 import world

    class Lazy(dict):
        def __missing__(self, name):
            if name == "superflat":
                w = world.generate_superflat()
            elif name == "plains":
                w = world.generate_plains(plains_columns)
            elif name == "gi_box":
                w = world.generate_gi_box(plains_columns)
            elif name == "city":
                w = world.generate_city()
            elif name == "orchard":
                w = world.generate_orchard(plains_columns)
            elif name == "empty":
                w = world.World()
            elif name == "sparse":
                rng = np.random.RandomState(5)
                w = world.World()
                idx = rng.randint(0, w.data.size, size=400)
                w.data[idx] = rng.randint(1, 100, size=400)
            else:
                raise KeyError(name)
            self[name] = w
            return w

    return Lazy()


@pytest.fixture(scope="session")
def oracle_dfs(worlds):
    from oracle import vxo

    class Lazy(dict):
        def __missing__(self, name):
            self[name] = vxo.df_build(worlds[name].data)
            return self[name]

    return Lazy()


@pytest.fixture(scope="session")
def scene_tables():
    from voxelpathtracer_b200 import assets, camera
    sun, moon, stronger, sunvis = camera.sun_moon_direction(50.0)
    return {
        "materials": assets.load_materials(),
        "blue_noise": assets.load_blue_noise(),
        "sky": assets.analytic_sky(16, sun),
        "shadow_noise": assets.load_shadow_noise(),
        "sun": sun, "moon": moon, "stronger": stronger, "sun_visibility": sunvis,
    }


@pytest.fixture(scope="session")
def oracles(worlds, oracle_dfs, scene_tables):
    from oracle import vxo

    class Lazy(dict):
        def __missing__(self, name):
            o = vxo.Oracle(worlds[name].data, oracle_dfs[name])
            o.set_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
            self[name] = o
            return o

    return Lazy()


@pytest.fixture(scope="session")
def renderer(scene_tables):
    """One CUDA handle for the whole GPU session (calls go through the C ABI via ctypes)."""
    import voxelpathtracer_b200 as vx
    r = vx.Renderer(0)
    r.load_scene_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
    yield r
    r.close()


@pytest.fixture(scope="session")
def golden_digests():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "oracle_digests.json")) as f:
        return json.load(f)
