"""ctypes front end shared by the two CPU checkers.

  CheckerLib("ref")    -> oracle/_ref/libg19ref.so : the UNMODIFIED reference
                          headers compiled behind oracle/ref_harness/ref_driver.cpp
  CheckerLib("oracle") -> oracle/libg19oracle.so   : this repo's plain-C
                          restatement (ref_restate.c) + path oracle (path_oracle.c)

Both export the same probe API (prefix g19ref_ / g19o_), so the tests can run
one body against either. TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C
import importlib
import os

import numpy as np

abi = importlib.import_module("2019global_b200.abi")

HERE = os.path.dirname(os.path.abspath(__file__))
PATHS = {"ref": os.path.join(HERE, "_ref", "libg19ref.so"), "oracle": os.path.join(HERE, "libg19oracle.so")}
PREFIX = {"ref": "g19ref_", "oracle": "g19o_"}


def available(which):
    return os.path.exists(PATHS[which])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CheckerLib:
    def __init__(self, which):
        self.which = which
        self.lib = C.CDLL(PATHS[which])
        self.pre = PREFIX[which]
        self.fn("scene_create").restype = C.c_void_p

    def fn(self, name):
        return getattr(self.lib, self.pre + name)

    def scene(self, root_min, root_max, descs):
        return CheckerScene(self, root_min, root_max, descs)


class CheckerScene:
    """An Octree plus its entities on one of the CPU checkers."""

    def __init__(self, lib, root_min, root_max, descs):
        self.lib = lib
        self.h = C.c_void_p(lib.fn("scene_create")(abi.d3(root_min), abi.d3(root_max)))
        self.descs = list(descs)
        for d in self.descs:
            idx = lib.fn("scene_add")(self.h, C.byref(d))
            assert idx >= 0

    def __del__(self):
        try:
            self.lib.fn("scene_destroy")(self.h)
        except Exception:
            pass

    def bbox(self, idx):
        out = (C.c_double * 6)()
        self.lib.fn("entity_bbox")(self.h, idx, out)
        return np.array(out[:])

    def triangles(self, idx, max_tris=256):
        out = np.zeros((max_tris, 9))
        n = self.lib.fn("entity_triangles")(self.h, idx, self.descs[idx].kind, _ptr(out), max_tris)
        return out[:n].copy()

    def triangle_derived(self, p9):
        p = np.ascontiguousarray(p9, dtype=np.float64)
        out = np.zeros(12)
        self.lib.fn("triangle_derived")(_ptr(p), _ptr(out))
        return out

    def intersect(self, idx, origins, dirs):
        o = np.ascontiguousarray(origins, dtype=np.float64)
        d = np.ascontiguousarray(dirs, dtype=np.float64)
        n = o.shape[0]
        hit = np.zeros(n, np.int32)
        pts = np.zeros((n, 3))
        nrm = np.zeros((n, 3))
        self.lib.fn("intersect")(self.h, idx, n, _ptr(o), _ptr(d), _ptr(hit), _ptr(pts), _ptr(nrm))
        return hit, pts, nrm

    def texcoord(self, idx, points):
        p = np.ascontiguousarray(points, dtype=np.float64)
        uv = np.zeros((p.shape[0], 2), np.int32)
        self.lib.fn("texcoord")(self.h, idx, p.shape[0], _ptr(p), _ptr(uv))
        return uv

    def shade(self, idx, o, d, light, point, normal, u, v):
        out = (C.c_double * 3)()
        self.lib.fn("shade")(self.h, idx, abi.d3(o), abi.d3(d), abi.d3(light), abi.d3(point), abi.d3(normal),
                             int(u), int(v), out)
        return np.array(out[:])

    def candidates(self, o, d, max_out=1 << 16):
        out = np.zeros(max_out, np.int32)
        n = self.lib.fn("candidates")(self.h, abi.d3(o), abi.d3(d), _ptr(out), max_out)
        return out[:min(n, max_out)].copy()

    def render(self, cam, light, w, h):
        """RayTracer::run verbatim (ref) / its restatement (oracle) -> (h,w,3) uint8."""
        rgb = np.zeros((h, w, 3), np.uint8)
        self.lib.fn("render")(self.h, C.byref(cam), abi.d3(light), w, h, _ptr(rgb))
        return rgb

    def trace(self, cam, light, w, h, y0=0, y1=None, want=("ids", "points", "normals", "rgb"), threads=1):
        y1 = h if y1 is None else y1
        ids = np.full((h, w), -2, np.int32) if "ids" in want else None
        pts = np.zeros((h, w, 3)) if "points" in want else None
        nrm = np.zeros((h, w, 3)) if "normals" in want else None
        rgb = np.zeros((h, w, 3), np.uint8) if "rgb" in want else None
        self.lib.fn("trace")(self.h, C.byref(cam), abi.d3(light), w, h, y0, y1, _ptr(ids), _ptr(pts), _ptr(nrm),
                             _ptr(rgb), threads)
        return {"ids": ids, "points": pts, "normals": nrm, "rgb": rgb}


def path_render(scene, cam, w, h, spp, max_depth, seed=0, window=None, threads=8):
    """FP64 brute-force path oracle (oracle/path_oracle.c) -> (radiance (h,w,3) float32, [extend, shadow])."""
    assert scene.lib.which == "oracle"
    x0, y0, x1, y1 = window if window else (0, 0, w, h)
    rad = np.zeros((h, w, 3), np.float32)
    segs = (C.c_uint64 * 2)()
    fn = scene.lib.lib.g19o_path_render
    fn(scene.h, C.byref(cam), w, h, spp, max_depth, C.c_uint32(seed), x0, y0, x1, y1, _ptr(rad), segs, threads)
    return rad, (int(segs[0]), int(segs[1]))
