"""Host side of the product (scene.cpp, builtin_scenes.cpp, C-ABI surface). No GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import mirror, zoo


def test_library_exports_every_declared_symbol(g19):
    L = g19.lib()
    header = open(os.path.join(os.path.dirname(g19.LIB_PATH), "..", "include", "g19.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|void|int64_t|const char\*)\s+(g19_[a-z_0-9]+)\s*\(", header, re.M)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), "include/g19.h declares %s but the library does not export it" % name
    assert set(g19.EXPORTS) <= set(declared)


def test_abi_struct_sizes(g19, abi, tmp_path):
    """The ctypes mirrors (2019global_b200/abi.py) against the C compiler's view of include/g19.h."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "g19.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %d\\n", '
                   'sizeof(g19_entity_desc), sizeof(g19_camera), sizeof(g19_params), sizeof(g19_stats), '
                   'offsetof(g19_entity_desc, shader_parameters), G19_ABI_VERSION); return 0; }\n')
    exe = str(tmp_path / "sizes")
    subprocess.run(["gcc", "-I", os.path.join(os.path.dirname(g19.LIB_PATH), "..", "include"), str(src), "-o", exe], check=True)
    desc, cam, params, stats, off, version = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(abi.EntityDesc) == desc == 224
    assert abi.EntityDesc.shader_parameters.offset == off
    assert ctypes.sizeof(abi.Camera) == cam == 56
    assert ctypes.sizeof(abi.Params) == params == 44
    assert ctypes.sizeof(abi.Stats) == stats
    assert abi.ABI_VERSION == version


def test_entity_geometry_matches_oracle_bits(g19, oracle):
    z = zoo(g19)
    chk = mirror(oracle, z)
    for i in range(len(z)):
        assert z.bbox(i).tobytes() == chk.bbox(i).tobytes(), i
        assert z.triangles(i).tobytes() == chk.triangles(i).tobytes(), i
    # a few randomised constructor arguments per composite kind
    rng = np.random.default_rng(3)
    sc = g19.Octree((-50,) * 3, (50,) * 3)
    for _ in range(10):
        p = rng.uniform(-3, 3, 3)
        sc.push_back(g19.ExpQuad(p, *rng.uniform(0.5, 4, 2), rng.uniform(0.1, 3.0), (1, 0, 0)))
        sc.push_back(g19.ExpCube(p, *rng.uniform(0.5, 4, 3), (1, 0, 0)))
        sc.push_back(g19.ExpCone(p, rng.uniform(-1, 1, 3), *rng.uniform(0.5, 4, 2), (1, 0, 0)))
        sc.push_back(g19.ExpSphere(p, rng.uniform(0.5, 3), (1, 0, 0)))
        sc.push_back(g19.ImpTriangle(*rng.uniform(-4, 4, (3, 3))))
    chk = mirror(oracle, sc)
    for i in range(len(sc)):
        assert sc.bbox(i).tobytes() == chk.bbox(i).tobytes(), i
        assert sc.triangles(i).tobytes() == chk.triangles(i).tobytes(), i


def test_default_scene_is_main_cpp_literal(g19, abi):
    sc, cam, light = g19.Octree.builtin(abi.SCENE_DEFAULT)
    kinds = [d.kind for d in sc.entities()]
    assert kinds == [abi.EXP_QUAD, abi.IMP_SPHERE, abi.IMP_SPHERE]  # main.cpp:46-48 push order
    assert tuple(cam.pos) == (-10, 0, 0) and tuple(cam.look_at) == (1, 0, 0) and cam.focal == 0.1
    assert light == (-10, 10, 10)
    # SURVEY hard part 5: the sphere's bbox is centred on the ORIGIN
    assert sc.bbox(1).tolist() == [-2, -2, -2, 2, 2, 2]


def test_builtin_scene_shapes(g19, abi):
    sc, cam, _ = g19.Octree.builtin(abi.SCENE_CORNELL, w=1920, h=1080)
    assert len(sc) == 14
    bs = [d.bsdf for d in sc.entities()]
    assert bs.count(abi.BSDF_EMITTER) == 2 and bs[-2:] == [abi.BSDF_DIFFUSE] * 2
    gl, _, _ = g19.Octree.builtin(abi.SCENE_CORNELL_GLASS, w=64, h=64)
    assert [d.bsdf for d in gl.entities()][-2:] == [abi.BSDF_MIRROR, abi.BSDF_GLASS]
    # pitch that re-centres the 16:9 frame: focal*sin(theta) = (w-h)/2 * 0.0002
    f = np.array(cam.look_at) - np.array(cam.pos)
    assert np.isclose(-f[2] * cam.focal, (1920 - 1080) / 2 * 0.0002)
    hf, _, _ = g19.Octree.builtin(abi.SCENE_HEIGHTFIELD, n=16, w=64, h=64)
    assert len(hf) == 2 * 16 * 16 + 2
    with pytest.raises(g19.G19Error):
        g19.Octree.builtin(abi.SCENE_HEIGHTFIELD, n=0)


def test_rejected_entity_is_numbered_but_absent(g19, abi):
    sc = g19.Octree((-1,) * 3, (1,) * 3)
    i0, ok0 = sc.push_back(g19.ImpTriangle((5, 5, 5), (6, 5, 5), (5, 6, 5)))
    i1, ok1 = sc.push_back(g19.ImpSphere((9, 9, 9), 0.5, (1, 1, 1)))  # bbox is origin-centred: accepted!
    assert (i0, ok0) == (0, False) and (i1, ok1) == (1, True)
    assert len(sc) == 2


def test_tile_map_matches_c_abi(g19):
    tiles = __import__("importlib").import_module("2019global_b200.tiles")
    for (w, h) in ((1920, 1080), (333, 217), (32, 32), (1, 1), (3840, 2160)):
        for world in (1, 2, 3, 8):
            seen = np.zeros(w * h, np.int32)
            for r in range(world):
                g = tiles.local_pixels(w, h, r, world)
                assert g.shape[0] == g19.engine.tile_pixels(w, h, r, world)
                seen[g[g >= 0]] += 1
            assert (seen == 1).all()  # every pixel owned exactly once


def test_no_gpu_fails_loudly(g19):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g19.G19Error) as e:
        g19.RayTracer(g19.Camera((-10, 0, 0)), (0, 0, 0))
    assert e.value.code == g19.abi.ERR_NO_DEVICE and "no CPU path" in str(e.value)


def test_stats_and_params_layout_matches_the_header(tmp_path):
    """The ctypes mirrors of g19_stats / g19_params / g19_camera (2019global_b200/abi.py) against the C header itself:
    sizes and the offsets of the last fields, as gcc lays them out."""
    import ctypes
    import os
    import subprocess
    import importlib
    abi = importlib.import_module("2019global_b200").abi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "g19.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(g19_stats), offsetof(g19_stats, shade_calls_folded), '
                   'offsetof(g19_stats, class_launches), sizeof(g19_params), sizeof(g19_camera)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(abi.Stats)
    assert int(out[1]) == abi.Stats.shade_calls_folded.offset
    assert int(out[2]) == abi.Stats.class_launches.offset
    assert int(out[3]) == ctypes.sizeof(abi.Params)
    assert int(out[4]) == ctypes.sizeof(abi.Camera)
