"""Run BASELINE.json's five configs on one GPU and print a table (evidence, not bench lines).

C1 default scene 500x500 REF | C2 Cornell 1080p 64spp d5 | C3 glass Cornell 1080p 64spp d12
C4 1M-triangle heightfield 1080p (spp reduced by --c4-spp) | C5 Cornell 4K (spp reduced by --c5-spp)
Also: REF-mode (the reference's own loop) frame times on the C2 and C4 scenes.
"""
import argparse
import importlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--c4-spp", type=int, default=64)
ap.add_argument("--c4-n", type=int, default=708)
ap.add_argument("--c5-spp", type=int, default=64)
a = ap.parse_args()
g19 = importlib.import_module("2019global_b200")
abi = g19.abi


def run(name, scene, n, w, h, mode, spp, depth, reps=2):
    t0 = time.time()
    sc, cam, light = g19.Octree.builtin(scene, n=n, w=w, h=h)
    t_build = time.time() - t0
    rt = g19.RayTracer(cam, light, device=0)
    t0 = time.time()
    rt.setScene(sc)
    t_up = time.time() - t0
    rt.start()
    best = None
    for _ in range(reps):
        out = rt.run(w, h, mode=mode, want=("rgb",), spp=spp, max_depth=depth, seed=0)
        st = rt.stats()
        best = st.render_ms if best is None else min(best, st.render_ms)
    ns = st.samples
    lit = float((out["rgb"].reshape(-1, 3).max(1) > 0).mean())
    print("| %s | %d | %dx%d | %s | %d | %d | %.2f | %.1f | %.2f | %.2f | %.0f%% | host build %.2fs, upload %.2fs |" % (
        name, len(sc), w, h, "REF" if mode == abi.MODE_REF else "PATH", spp, depth, best, ns / best / 1e3,
        st.extend_segments / max(1, ns), st.shadow_segments / max(1, ns), 100 * lit, t_build, t_up))
    sys.stdout.flush()


print("| config | entities | size | mode | spp | depth | ms/frame | Msamples/s | extend/sample | shadow/sample | non-black | notes |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
run("C1 default scene", abi.SCENE_DEFAULT, 0, 500, 500, abi.MODE_REF, 1, 1)
run("C1 default scene", abi.SCENE_DEFAULT, 0, 1920, 1080, abi.MODE_REF, 1, 1)
run("C2 Cornell (REF loop)", abi.SCENE_CORNELL, 0, 1920, 1080, abi.MODE_REF, 1, 1)
run("C2 Cornell", abi.SCENE_CORNELL, 0, 1920, 1080, abi.MODE_PATH, 64, 5)
run("C3 Cornell glass+mirror", abi.SCENE_CORNELL_GLASS, 0, 1920, 1080, abi.MODE_PATH, 64, 12)
run("C4 heightfield room n=%d (REF loop)" % a.c4_n, abi.SCENE_HEIGHTFIELD_ROOM, a.c4_n, 1920, 1080, abi.MODE_REF, 1, 1, reps=2)
run("C4 heightfield room n=%d" % a.c4_n, abi.SCENE_HEIGHTFIELD_ROOM, a.c4_n, 1920, 1080, abi.MODE_PATH, a.c4_spp, 5)
run("   open heightfield n=%d (round 1's C4 scene)" % a.c4_n, abi.SCENE_HEIGHTFIELD, a.c4_n, 1920, 1080, abi.MODE_PATH, a.c4_spp, 5)
run("C5 Cornell 4K", abi.SCENE_CORNELL, 0, 3840, 2160, abi.MODE_PATH, a.c5_spp, 8)
