"""The PATH-mode linear octree built on the GPU (csrc/tree_build.cu) against
the host builder (csrc/path.cu) that defines it: node records and leaf lists must be IDENTICAL, and
so must the rendered frames. Reference counterpart: Octree::push_back / Node::partition
(reference include/octree.h:20-43,75-129)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tree(g19, abi, sc, cam, light, how):
    os.environ["G19_TREE_BUILD"] = how
    try:
        rt = g19.RayTracer(cam, light)
        rt.setScene(sc)
    finally:
        del os.environ["G19_TREE_BUILD"]
    rt.start()
    nodes, index = rt.path_tree()
    return rt, nodes, index


def _check_tree(nodes, index, n_prims_min):
    leaf = (nodes[:, 1] >> 31) == 1
    cnt = nodes[:, 1] & 0x7fffffff
    assert int(cnt[leaf].sum()) == index.size           # leaf lists tile the index array
    inner = ~leaf
    assert (nodes[inner, 1] == 0).all()
    first_children = np.sort(nodes[inner, 0])
    assert np.array_equal(first_children, 1 + 8 * np.arange(inner.sum()))  # breadth first, complete blocks of 8
    assert np.unique(index).size >= n_prims_min           # every primitive is reachable


@pytest.mark.parametrize("n", [24, 96, 300])
def test_device_tree_equals_host_tree(g19, abi, n):
    w, h = 96, 54
    sc, cam, light = g19.Octree.builtin(abi.SCENE_HEIGHTFIELD, n=n, w=w, h=h)
    rt_h, nodes_h, index_h = _tree(g19, abi, sc, cam, light, "host")
    rt_d, nodes_d, index_d = _tree(g19, abi, sc, cam, light, "device")
    print("n=%d: %d nodes, %d leaf references" % (n, nodes_h.shape[0], index_h.size))
    assert nodes_h.shape[0] > 1
    _check_tree(nodes_h, index_h, 2 * n * n * 0.9)
    assert np.array_equal(nodes_h, nodes_d)
    assert np.array_equal(index_h, index_d)
    a = rt_h.run(w, h, mode=abi.MODE_PATH, want=("radiance",), spp=2, max_depth=3, seed=4)["radiance"]
    b = rt_d.run(w, h, mode=abi.MODE_PATH, want=("radiance",), spp=2, max_depth=3, seed=4)["radiance"]
    assert a.tobytes() == b.tobytes() and a.max() > 0


def test_device_build_small_and_empty_scenes(g19, abi):
    """Forced device build of scenes the host builder would keep flat, and of an empty scene."""
    sc, cam, light = g19.Octree.builtin(abi.SCENE_CORNELL, w=64, h=36)
    rt_h, nodes_h, index_h = _tree(g19, abi, sc, cam, light, "host")
    rt_d, nodes_d, index_d = _tree(g19, abi, sc, cam, light, "device")
    assert np.array_equal(nodes_h, nodes_d) and np.array_equal(index_h, index_d)
    a = rt_h.run(64, 36, mode=abi.MODE_PATH, want=("radiance",), spp=4, max_depth=4)["radiance"]
    b = rt_d.run(64, 36, mode=abi.MODE_PATH, want=("radiance",), spp=4, max_depth=4)["radiance"]
    assert a.tobytes() == b.tobytes()
    empty = g19.Octree((-1,) * 3, (1,) * 3)
    rt_e, nodes_e, index_e = _tree(g19, abi, empty, cam, light, "device")
    assert nodes_e.shape[0] == 1 and index_e.size == 0
    assert (rt_e.run(40, 24, mode=abi.MODE_PATH, want=("radiance",), spp=2, max_depth=3)["radiance"] == 0).all()


def test_upload_threads_do_not_change_the_scene(g19, abi):
    """Scene upload extracts primitives on several host threads (a range of entities each, joined in entity order with
    the material folding applied across the seams): tree, leaf references and radiance must not depend on their number.
    The zoo mixes materials, spheres and triangle composites; the heightfield is the many-entity case; sort_rays on/off
    (tree scenes walk each bounce's rays in origin-cell order) must agree bit for bit as well."""
    w, h = 96, 54
    from util import zoo
    scenes = [g19.Octree.builtin(abi.SCENE_HEIGHTFIELD_ROOM, n=40, w=w, h=h)]
    cam = g19.Camera((-10, 0, 0), (1, 0, 0), 0.1)
    scenes.append((zoo(g19), cam, (-10, 10, 10)))
    for sc, cam, light in scenes:
        outs = []
        for threads, sort in ((1, 1), (3, 1), (7, 0), (16, 1)):
            rt = g19.RayTracer(cam, light)
            rt.tune("upload_threads", threads)
            rt.tune("sort_rays", sort)
            rt.tune("tree_build", "host")
            rt.setScene(sc)
            rt.start()
            nodes, index = rt.path_tree()
            rad = rt.run(w, h, mode=abi.MODE_PATH, want=("radiance",), spp=3, max_depth=4, seed=9)["radiance"]
            outs.append((nodes, index, rad))
        for nodes, index, rad in outs[1:]:
            assert np.array_equal(nodes, outs[0][0]) and np.array_equal(index, outs[0][1])
            assert rad.tobytes() == outs[0][2].tobytes()
