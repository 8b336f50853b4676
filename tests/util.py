"""Shared helpers for the parity tests."""
import contextlib
import math
import os

import numpy as np


@contextlib.contextmanager
def quiet_stdout():
    """The reference prints from ExpRectangle::getTextureCoord (entities.h:362); silence fd 1."""
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def mirror(checker, octree):
    """The same Octree + entities on a CPU checker, from the product scene's own descriptors."""
    return checker.scene(octree.min, octree.max, octree.entities())


def zoo(g19):
    """One entity of every reference class, in a root that subdivides."""
    E = g19
    sc = g19.Octree((-20,) * 3, (20,) * 3)
    descs = [
        E.ImpSphere((3, 4, 4), 2, (1, 0, 0)),
        E.ImpTriangle((3, -2, 1), (3, 2, -1), (3, -2, -1), (0, 1, 0)),
        E.ExpRectangle((0, 0, 0), (3, 3, 3), (3, 3, 0), (1, 1, 0)),
        E.ExpBox((1, 1, 1), (3, 4, 5), (0, 1, 1)),
        E.ExpSphere((-2, 0, 0), 2, (0, 1, 0)),
        E.ExpQuad((0.5, 0.2, 0.3), 2, 3, 0.7, (1, 2, 3)),
        E.ExpCube((0.3, -0.4, 1), 2, 2.5, 1.5, (1, 0, 0)),
        E.ExpCone((0, 0, 2), (-1, 1, -3), 5, 3, (1, 1, 0)),
        E.ImpSphere((-3, -4, -4), 1.5, (0, 0, 1)),
    ]
    for d in descs:
        sc.push_back(d)
    return sc


def probe_rays(n, seed=0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-12, 12, (n, 3))
    tgt = rng.uniform(-4, 4, (n, 3))
    d = tgt - o
    k = min(50, n // 8)
    d[:k] = [1, 0, 0]
    d[k:2 * k] = [0, 1, 0]
    d[2 * k:3 * k] = [0, 0, 1]
    d[3 * k:4 * k] = [0, 0, -1]
    return o, d


def rel_rmse(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return math.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-12)
