"""CPU-only properties of the path oracle (oracle/path_oracle.c): the definition must at least be
self-consistent before the GPU is compared against it."""
import numpy as np

from oracle import binding
from util import mirror


def test_deterministic_and_thread_invariant(g19, abi, oracle):
    w, h = 40, 24
    sc, cam, _ = g19.Octree.builtin(abi.SCENE_CORNELL_GLASS, w=w, h=h)
    chk = mirror(oracle, sc)
    a, sa = binding.path_render(chk, cam, w, h, 4, 8, seed=3, threads=1)
    b, sb = binding.path_render(chk, cam, w, h, 4, 8, seed=3, threads=5)
    assert a.tobytes() == b.tobytes() and sa == sb
    c, _ = binding.path_render(chk, cam, w, h, 4, 8, seed=4, threads=5)
    assert a.tobytes() != c.tobytes()
    # a window renders the same pixels as the full frame (the bench's bounded crop)
    win, _ = binding.path_render(chk, cam, w, h, 4, 8, seed=3, window=(8, 4, 24, 20), threads=2)
    assert win[4:20, 8:24].tobytes() == a[4:20, 8:24].tobytes()
    assert win[:4].sum() == 0


def test_depth_one_is_direct_lighting(g19, abi, oracle):
    """max_depth counts segments of the CAMERA path: depth 1 = emitters seen directly + one
    next-event sample at the first diffuse hit, nothing indirect."""
    w, h = 64, 36
    sc, cam, _ = g19.Octree.builtin(abi.SCENE_CORNELL, w=w, h=h)
    rad, segs = binding.path_render(mirror(oracle, sc), cam, w, h, 4, 1, seed=0)
    assert segs[0] == w * h * 4 and 0 < segs[1] <= segs[0]
    assert np.isclose(rad.max(), 17.0)  # the ceiling light, seen directly
    deeper, _ = binding.path_render(mirror(oracle, sc), cam, w, h, 4, 5, seed=0)
    assert deeper.mean() > rad.mean()  # indirect light only adds


def test_radiance_is_linear_in_emission(g19, abi, oracle):
    """Size-independent property: doubling every emitter doubles every pixel, bit for bit."""
    w, h = 32, 18
    sc, cam, _ = g19.Octree.builtin(abi.SCENE_CORNELL, w=w, h=h)
    descs = sc.entities()
    a, _ = binding.path_render(oracle.scene(sc.min, sc.max, descs), cam, w, h, 8, 5, seed=1)
    for d in descs:
        for k in range(3):
            d.emission[k] *= 2.0
    b, _ = binding.path_render(oracle.scene(sc.min, sc.max, descs), cam, w, h, 8, 5, seed=1)
    assert (2.0 * a.astype(np.float64)).astype(np.float32).tobytes() == b.tobytes()
    assert a.mean() > 0.05


def test_energy_is_bounded(g19, abi, oracle):
    """Closed white-ish room: radiance stays finite and below the emitter's."""
    w, h = 32, 18
    sc, cam, _ = g19.Octree.builtin(abi.SCENE_CORNELL, w=w, h=h)
    rad, _ = binding.path_render(mirror(oracle, sc), cam, w, h, 16, 8, seed=2)
    assert np.isfinite(rad).all() and rad.min() >= 0 and rad.max() <= 17.0 * 1.0001


def test_grid_accelerator_is_bit_identical_to_brute_force(g19, abi, oracle):
    """Large scenes (bench.py's C4 windows) go through a uniform grid inside the oracle; it must return exactly what
    the loop over every primitive returns -- radiance, segment counts and the primary-hit AOV, bit for bit."""
    from util import zoo
    cases = [(g19.Octree.builtin(abi.SCENE_CORNELL_GLASS, w=64, h=36), 64, 36, 8, 9),
             (g19.Octree.builtin(abi.SCENE_HEIGHTFIELD, n=40, w=64, h=36), 64, 36, 4, 5),
             (g19.Octree.builtin(abi.SCENE_HEIGHTFIELD_ROOM, n=40, w=64, h=36), 64, 36, 4, 5),
             ((zoo(g19), g19.Camera((-10, 0, 0), (1, 0, 0), 0.02), None), 48, 48, 4, 4)]
    try:
        for (sc, cam, _), w, h, spp, depth in cases:
            chk = mirror(oracle, sc)
            out = []
            for mode in (1, 2):  # brute force, grid always
                oracle.lib.g19o_path_set_accel(mode)
                rad, segs = binding.path_render(chk, cam, w, h, spp, depth, seed=3, threads=4)
                ids, pts, nrm = binding.path_primary(chk, cam, w, h)
                out.append((rad.tobytes(), segs, ids.tobytes(), pts.tobytes(), nrm.tobytes()))
            assert out[0] == out[1]
    finally:
        oracle.lib.g19o_path_set_accel(0)


def test_oracle_scene_generator_equals_the_products(g19, abi):
    """bench.py's CPU legs build their scenes with oracle/oracle_scenes.c (so the reference arm never loads the
    product library): every descriptor, the camera and the light must equal the product generator's."""
    for which, n, w, h in ((abi.SCENE_DEFAULT, 0, 500, 500), (abi.SCENE_CORNELL, 0, 1920, 1080), (abi.SCENE_CORNELL_GLASS, 0, 640, 360),
                           (abi.SCENE_HEIGHTFIELD, 24, 320, 200), (abi.SCENE_HEIGHTFIELD_ROOM, 24, 3840, 2160)):
        sc, cam, light = g19.Octree.builtin(which, n=n, w=w, h=h)
        descs, cam2, light2 = binding.builtin_descs(which, n, w, h)
        mine = sc.entities()
        assert len(mine) == len(descs)
        assert all(bytes(a) == bytes(b) for a, b in zip(mine, descs))
        assert bytes(cam) == bytes(cam2) and tuple(light) == tuple(light2)
