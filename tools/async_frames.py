"""Frames enqueued back to back through the device-pointer entry point (no host synchronisation in between), timed with
CUDA events: what bench.py's timed region does, for one GPU or for one rank's share of the tiles (--world N)."""
import argparse
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="CORNELL")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--depth", type=int, default=5)
ap.add_argument("--frames", type=int, default=20)
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--tune", action="append", default=[])
a = ap.parse_args()
g19 = importlib.import_module("2019global_b200")
abi = g19.abi
sc, cam, light = g19.Octree.builtin(getattr(abi, "SCENE_" + a.scene), n=a.n, w=a.w, h=a.h)
rt = g19.RayTracer(cam, light, device=0)
for kv in a.tune:
    k, v = kv.split("=", 1)
    rt.tune(k, v)
rt.setScene(sc)
rt.start()
buf = torch.zeros(a.h * a.w * 3, dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
p = rt.params(a.w, a.h, mode=abi.MODE_PATH, spp=a.spp, max_depth=a.depth, seed=0, rank=0, world=a.world)
for _ in range(4):
    rt.run_device(p, d_rad=buf.data_ptr(), stream=stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.frames):
    rt.run_device(p, d_rad=buf.data_ptr(), stream=stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.frames
share = a.w * a.h * a.spp / a.world
print("async: %s %dx%d spp %d depth %d rank 0 of %d %s: %.3f ms per frame, %.1f Msamples/s" % (
    a.scene, a.w, a.h, a.spp, a.depth, a.world, " ".join(a.tune), ms, share / ms / 1e3))
