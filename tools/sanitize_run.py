"""Tiny run of every kernel family (REF + probes, PATH flat / tree / glass, tiles) for compute-sanitizer."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import probe_rays, zoo  # noqa: E402

g19 = importlib.import_module("2019global_b200")
abi = g19.abi
z = zoo(g19)
rt = g19.RayTracer(g19.Camera((-10, 0, 0), (1, 0, 0), 0.1), (-10, 10, 10), device=0)
rt.setScene(z)
rt.start()
rt.run(97, 61, want=("rgb", "ids", "radiance"))
for i in range(len(z)):
    o, d = probe_rays(256, seed=i)
    rt.probe_intersect(i, o, d)
rt.probe_candidates((-10, 0.3, 0.2), (1, 0.01, 0.02))
for which, n, depth in ((abi.SCENE_CORNELL_GLASS, 0, 8), (abi.SCENE_HEIGHTFIELD, 40, 4), (abi.SCENE_CORNELL, 0, 5)):
    sc, cam, light = g19.Octree.builtin(which, n=n, w=97, h=61)
    rt.camera, rt.light = cam, light
    rt.setScene(sc)
    rt.run(97, 61, mode=abi.MODE_PATH, want=("rgb", "radiance"), spp=5, max_depth=depth, spp_per_pass=2)
    out = {"rgb": np.zeros((61, 97, 3), np.uint8)}
    for r in range(3):
        rt.run(97, 61, mode=abi.MODE_PATH, want=("rgb",), out=out, spp=2, max_depth=3, rank=r, world=3)
print("sanitize_run ok")
