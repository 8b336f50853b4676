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
