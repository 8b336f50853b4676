"""Short single-GPU run of the bench workload for ncu (one frame, fewer spp): exits 0 on success."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--depth", type=int, default=5)
ap.add_argument("--scene", default="CORNELL")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--mode", default="PATH")
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--profile", type=int, default=0)
ap.add_argument("--spp-per-pass", type=int, default=0)
ap.add_argument("--pixels-per-pass", type=int, default=0)
ap.add_argument("--tune", action="append", default=[], help="key=value for g19_tune (repeatable)")
ap.add_argument("--world", type=int, default=1, help="render only the tiles rank --rank of --world ranks owns (one GPU emulating one rank of N)")
ap.add_argument("--rank", type=int, default=0)
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
for _ in range(a.frames):
    out = rt.run(a.w, a.h, mode=getattr(abi, "MODE_" + a.mode), want=("rgb",), spp=a.spp, max_depth=a.depth, seed=0, profile=a.profile, spp_per_pass=a.spp_per_pass, pixels_per_pass=a.pixels_per_pass, rank=a.rank, world=a.world)
st = rt.stats()
if a.world > 1:
    a.tune = a.tune + ["rank %d of %d" % (a.rank, a.world)]
print("ok: %s %s n=%d %dx%d spp %d depth %d %s: %.2f ms, %.1f Msamples/s, %d samples, %d extend, %d shadow segments, %d launches, node/prim tests %d/%d" % (
    a.mode, a.scene, a.n, a.w, a.h, a.spp, a.depth, " ".join(a.tune), st.render_ms, st.samples / max(st.render_ms, 1e-9) / 1e3, st.samples,
    st.extend_segments, st.shadow_segments, st.kernel_launches, st.node_tests, st.prim_tests))
if a.profile:
    print("  class ms:", {abi.CLASS_NAMES[k]: round(st.class_ms[k], 2) for k in range(7) if st.class_launches[k]},
          "launches:", {abi.CLASS_NAMES[k]: st.class_launches[k] for k in range(7) if st.class_launches[k]})
