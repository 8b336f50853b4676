"""Closed-form radiance checks of G19_MODE_PATH -- for BOTH the oracle (oracle/path_oracle.c) and the CUDA tracer.

PATH mode's transport (next-event estimation, cosine-weighted bounces, Fresnel split, depth limit) has no
counterpart in the reference (reference include/raytracer.h:32-86 casts one ray and shades it directly), and the
oracle that defines it was written next to the kernels: a shared misconception -- a wrong NEE weight, a wrong cosine
pdf, a wrong Fresnel term -- would pass every oracle-vs-GPU test. These scenes have answers that come from radiometry,
not from either implementation:

  furnace            a convex lambertian sphere (albedo rho) inside a closed box emitting E from every face: the
                     sphere's radiance is rho * E at any depth (it sees E over its whole hemisphere and never itself);
                     pixels that see the box show E exactly
  integrating sphere a small two-sided emitter (area A, radiance Le) at the centre of a closed lambertian sphere seen
                     from inside (radius R, albedo rho): every wall element lights every other one equally, so after
                     the direct term E0(x) = Le A |cos| / R^2 each further bounce adds the UNIFORM irradiance
                     rho^(k-1) * Le A / (2 R^2): L(x) = rho/pi * [E0(x) + Le A/(2 R^2) * sum_{k=2..depth} rho^(k-1)].
                     Checks the NEE weight, the cosine-weighted bounce (pdf cancels cos/pi), the throughput product,
                     the meaning of max_depth (segments per path) and shading a sphere from inside
  fresnel            a glass ball (ior n) in front of the camera, an emitter behind the camera, black elsewhere: at
                     normal incidence the light coming back is Le * [F + (1-F)^2 F (1 + F^2 + ...)] = Le * 2F/(1+F),
                     F = ((n-1)/(n+1))^2; truncated at depth 2 it is Le * F alone
"""
import math

import numpy as np
import pytest

from util import mirror

E_BOX = (1.0, 0.5, 0.25)
RHO = (0.8, 0.5, 0.3)


def _quad(g19, sc, a, b, c, d, **kw):
    sc.push_back(g19.ImpTriangle(a, b, c, **kw))
    sc.push_back(g19.ImpTriangle(a, c, d, **kw))


def _box(g19, sc, h, **kw):
    """closed axis-aligned box [-h, h]^3 as 12 triangles"""
    for axis in range(3):
        for s in (-h, h):
            u, v = [(1, 2), (0, 2), (0, 1)][axis]
            corners = []
            for cu, cv in ((-h, -h), (h, -h), (h, h), (-h, h)):
                p = [0.0, 0.0, 0.0]
                p[axis], p[u], p[v] = s, cu, cv
                corners.append(tuple(p))
            _quad(g19, sc, *corners, **kw)


def furnace_scene(g19, abi):
    sc = g19.Octree((-20,) * 3, (20,) * 3)
    _box(g19, sc, 12.0, color=(1, 1, 1), bsdf=abi.BSDF_EMITTER, emission=E_BOX)
    sc.push_back(g19.ImpSphere((0, 0, 0), 2.0, RHO))
    cam = g19.Camera((-10, 0, 0), (1, 0, 0), 0.02)  # 64 px wide (pitch 0.0002): the ball (angular radius 0.2) is ~20 px in radius
    return sc, cam


def check_furnace(rad, ids):
    ball = ids == 12
    assert ball.sum() > 500 and (~ball).sum() > 500
    inner = ball.copy()  # a pixel's samples are jittered over [x, x+1) x [y, y+1): keep pixels whose neighbours are ball too
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            inner &= np.roll(np.roll(ball, dy, 0), dx, 1)
    got = rad[inner].astype(np.float64).mean(0)
    exp = np.array(RHO) * np.array(E_BOX)
    assert np.allclose(got, exp, rtol=0.015), (got, exp)
    # what is not the ball is the box, seen directly: E, every sample, exactly (silhouette pixels mix the two)
    far = ~ball
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            far &= ~np.roll(np.roll(ball, dy, 0), dx, 1)
    assert np.allclose(rad[far], np.array(E_BOX, np.float32), rtol=1e-6)


R_INT, RHO_INT, LE_INT, HALF_INT = 5.0, 0.7, 400.0, 0.05


def integrating_sphere_scene(g19, abi):
    sc = g19.Octree((-20,) * 3, (20,) * 3)
    sc.push_back(g19.ImpSphere((0, 0, 0), R_INT, (RHO_INT,) * 3))
    a = HALF_INT  # emitter in the plane x = 0, normal along x: the camera (on the -x side) sees it edge... face on
    _quad(g19, sc, (0, -a, -a), (0, a, -a), (0, a, a), (0, -a, a), color=(1, 1, 1), bsdf=abi.BSDF_EMITTER, emission=(LE_INT,) * 3)
    cam = g19.Camera((-3, 0, 0), (1, 0, 0), 0.0064 / 0.6)  # 64 px wide, half-angle ~31 degrees
    return sc, cam


def expected_integrating(pts, depth):
    """L at wall points pts (n,3) for max_depth = depth segments."""
    area = (2 * HALF_INT) ** 2
    cos_l = np.abs(pts[:, 0]) / np.linalg.norm(pts, axis=1)
    e0 = LE_INT * area * cos_l / R_INT ** 2
    uniform = LE_INT * area / (2 * R_INT ** 2) * sum(RHO_INT ** (k - 1) for k in range(2, depth + 1))
    return RHO_INT / math.pi * (e0 + uniform)


def check_integrating(rad, ids, pts, depth):
    wall = ids == 0
    # keep away from the emitter's silhouette and from its shadow-free but finite-size penumbra effects: use wall pixels
    # whose 3x3 neighbourhood is wall
    m = wall.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            m &= np.roll(np.roll(wall, dy, 0), dx, 1)
    m[0, :] = m[-1, :] = False
    m[:, 0] = m[:, -1] = False
    assert m.sum() > 2000
    exp = expected_integrating(pts[m], depth)
    got = rad[m].astype(np.float64).mean(1)
    # image-wide: the mean is within 1 %; structure: four quadrant bands by |cos| follow the closed form within 2 %
    assert abs(got.mean() / exp.mean() - 1) < 0.01, (got.mean(), exp.mean())
    order = np.argsort(exp)
    for part in np.array_split(order, 4):
        assert abs(got[part].mean() / exp[part].mean() - 1) < 0.02, (got[part].mean(), exp[part].mean())


IOR, LE_F = 1.5, 3.0


def fresnel_scene(g19, abi):
    sc = g19.Octree((-20,) * 3, (20,) * 3)
    _quad(g19, sc, (-12, -6, -6), (-12, 6, -6), (-12, 6, 6), (-12, -6, 6), color=(1, 1, 1), bsdf=abi.BSDF_EMITTER,
          emission=(LE_F,) * 3)
    sc.push_back(g19.ImpSphere((0, 0, 0), 1.0, (1, 1, 1), bsdf=abi.BSDF_GLASS, ior=IOR))
    cam = g19.Camera((-10, 0, 0), (1, 0, 0), 0.1)
    return sc, cam


def check_fresnel(rad, depth):
    c = rad.shape[1] // 2
    patch = rad[c - 4:c + 5, c - 4:c + 5].astype(np.float64).mean()
    f = ((IOR - 1) / (IOR + 1)) ** 2
    exp = LE_F * (f if depth < 4 else 2 * f / (1 + f))
    assert abs(patch / exp - 1) < 0.03, (patch, exp, depth)


# ---- the oracle (CPU) ---------------------------------------------------------------------------------------------
def _oracle_render(oracle, sc, cam, w, h, spp, depth, seed=7):
    from oracle import binding
    chk = mirror(oracle, sc)
    rad, _ = binding.path_render(chk, cam, w, h, spp, depth, seed=seed, threads=8)
    ids, pts, _ = binding.path_primary(chk, cam, w, h)
    return rad, ids, pts


@pytest.mark.parametrize("depth", [1, 2, 5])
def test_oracle_furnace(g19, abi, oracle, depth):
    sc, cam = furnace_scene(g19, abi)
    rad, ids, _ = _oracle_render(oracle, sc, cam, 64, 64, 64, depth)
    check_furnace(rad, ids)


@pytest.mark.parametrize("depth", [1, 2, 3, 6])
def test_oracle_integrating_sphere(g19, abi, oracle, depth):
    sc, cam = integrating_sphere_scene(g19, abi)
    rad, ids, pts = _oracle_render(oracle, sc, cam, 64, 64, 48, depth)
    check_integrating(rad, ids, pts, depth)


@pytest.mark.parametrize("depth", [2, 12])
def test_oracle_fresnel_normal_incidence(g19, abi, oracle, depth):
    sc, cam = fresnel_scene(g19, abi)
    rad, _, _ = _oracle_render(oracle, sc, cam, 64, 64, 2048, depth)
    check_fresnel(rad, depth)


# ---- the CUDA tracer ------------------------------------------------------------------------------------------------
def _gpu_render(g19, abi, sc, cam, w, h, spp, depth, seed=7):
    rt = g19.RayTracer(cam, (0, 0, 0))
    rt.setScene(sc)
    rt.start()
    out = rt.run(w, h, mode=abi.MODE_PATH, want=("radiance", "ids"), spp=spp, max_depth=depth, seed=seed)
    return out["radiance"], out["ids"]


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [1, 2, 5])
def test_gpu_furnace(g19, abi, depth):
    sc, cam = furnace_scene(g19, abi)
    rad, ids = _gpu_render(g19, abi, sc, cam, 64, 64, 256, depth)
    check_furnace(rad, ids)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [1, 2, 3, 6])
def test_gpu_integrating_sphere(g19, abi, oracle, depth):
    from oracle import binding
    sc, cam = integrating_sphere_scene(g19, abi)
    rad, ids = _gpu_render(g19, abi, sc, cam, 64, 64, 256, depth)
    _, pts, _ = binding.path_primary(mirror(oracle, sc), cam, 64, 64)  # wall points of the pixels (geometry only)
    check_integrating(rad, ids, pts, depth)


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [2, 12])
def test_gpu_fresnel_normal_incidence(g19, abi, depth):
    sc, cam = fresnel_scene(g19, abi)
    rad, _ = _gpu_render(g19, abi, sc, cam, 64, 64, 8192, depth)
    check_fresnel(rad, depth)
