"""voxelpathtracer_b200 — B200-native distance-field build + DF-accelerated voxel traversal (sm_100a).

The product is libvxpt.so (CUDA kernels behind the C ABI of include/vxpt.h).  This package holds the build
recipe, the ctypes binding and the host-side mirror of the reference's World / camera / table interfaces.
There is no CPU path: importing works anywhere, but creating a Renderer needs the built library and a GPU.
"""
from . import abi, assets, camera, denoise, world  # noqa: F401
from .renderer import MultiRenderer, Renderer, diffuse_params, material_params, primary_params, reflection_params, shadow_params  # noqa: F401

__version__ = "0.1.0"
