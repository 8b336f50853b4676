"""2019global_b200 -- the B200-native radiance loop of preon7/2019global.

The package name starts with a digit, so import it with
    g19 = importlib.import_module("2019global_b200")
The compute path is lib2019global_b200.so (hand-written sm_100a CUDA behind the
C ABI of include/g19.h); there is no CPU fallback.
"""
from . import abi  # noqa: F401
from .engine import (Camera, ExpBox, ExpCone, ExpCube, ExpQuad, ExpRectangle, ExpSphere, G19Error, ImpSphere,  # noqa: F401
                     ImpTriangle, Octree, RayTracer, lib, EXPORTS, LIB_PATH)
