/* path_oracle.c -- FP64, brute-force CPU definition of G19_MODE_PATH.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_scene.h).
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference casts one primary ray per
 * pixel and shades it directly (reference include/raytracer.h:32-86,
 * include/material.h:48-62); it has no spp loop, no jitter, no RNG, no bounce
 * sampling, no area light, no mirror/glass.
 * This file is therefore the DEFINITION the CUDA wavefront tracer
 * (2019global_b200/csrc/path_kernels.cu) is checked against, not a restatement:
 *   - camera: the reference pinhole (raytracer.h:26-30,41) + uniform jitter
 *   - geometry: the primitives each entity tests in REF mode (spheres analytic,
 *     everything else its triangle list), float-rounded vertices, nearest hit
 *     with t > 0, found by testing EVERY primitive (no octree: this also checks
 *     the GPU's tree)
 *   - lambert / mirror / dielectric, triangle emitters, next-event estimation,
 *     emission on camera rays and after specular bounces only
 *   - Philox4x32-7 keyed on (pixel, sample, bounce, stream); the integer
 *     stream is bit-identical to the device's, discrete choices that depend on
 *     a float product are made in float like the device
 * What IS pinned: the camera basis comes from the same expression as
 * ref_restate.c (byte-equal to the compiled reference), and a max_depth=1 image
 * of a scene without emitters is black wherever REF mode reports a miss.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_scene.h"

#define RAY_EPS 1.0e-3
#define PI_D 3.14159265358979323846

typedef struct { double x, y, z; } v3;
static v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vmul(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 vmulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static v3 vnorm(v3 a) { return vmul(a, 1.0 / sqrt(vdot(a, a))); }

typedef struct {
    int is_tri;
    v3 v0, e1, e2, n; /* triangle */
    v3 c; double r;   /* sphere */
    int bsdf;
    v3 albedo, emission;
    double ior;
    int entity;
} p_prim;

typedef struct { v3 v0, e1, e2, n, emission; double area; } p_light;

/* Uniform grid over the primitives' bounding boxes (large scenes only). It is an ACCELERATOR of this oracle's own
 * brute-force loops, not a restatement of anything: closest() / occluded() through it return exactly what the
 * loops over every primitive return (tests/test_path_oracle.py checks bit-equality of whole frames), and it shares
 * nothing with the octree the CUDA tracer walks (a 3-D DDA over equal cells here, a parametric octree walk there). */
typedef struct {
    int on;
    double lo[3], cell[3], inv_cell[3];
    int res[3];
    int* start; /* CSR: cell c holds refs[start[c] .. start[c+1]) */
    int* refs;
} p_grid;

typedef struct {
    int n_prims, n_lights;
    p_prim* prims;
    p_light* lights;
    p_grid grid;
} p_scene;

static int g_accel = 0; /* 0 auto (grid from 512 primitives), 1 brute force, 2 grid always */
void g19o_path_set_accel(int mode) { g_accel = mode; }

static double clamp01(double v) { return v < 0 ? 0 : (v > 1 ? 1 : v); }
static v3 fround(d3 p) { return V((double)(float)p.x, (double)(float)p.y, (double)(float)p.z); }

static void prim_box(const p_prim* p, double lo[3], double hi[3]) {
    if (p->is_tri) {
        v3 a = p->v0, b = vadd(p->v0, p->e1), c = vadd(p->v0, p->e2);
        lo[0] = fmin(a.x, fmin(b.x, c.x)); hi[0] = fmax(a.x, fmax(b.x, c.x));
        lo[1] = fmin(a.y, fmin(b.y, c.y)); hi[1] = fmax(a.y, fmax(b.y, c.y));
        lo[2] = fmin(a.z, fmin(b.z, c.z)); hi[2] = fmax(a.z, fmax(b.z, c.z));
    } else {
        lo[0] = p->c.x - p->r; hi[0] = p->c.x + p->r;
        lo[1] = p->c.y - p->r; hi[1] = p->c.y + p->r;
        lo[2] = p->c.z - p->r; hi[2] = p->c.z + p->r;
    }
}

static void cell_range(const p_grid* g, const double lo[3], const double hi[3], int c0[3], int c1[3]) {
    for (int k = 0; k < 3; ++k) { /* boxes are padded by a hundredth of a cell: a hit on a cell wall is listed on both sides */
        double a = (lo[k] - g->lo[k]) * g->inv_cell[k] - 0.01, b = (hi[k] - g->lo[k]) * g->inv_cell[k] + 0.01;
        int ia = (int)floor(a), ib = (int)floor(b);
        c0[k] = ia < 0 ? 0 : (ia >= g->res[k] ? g->res[k] - 1 : ia);
        c1[k] = ib < 0 ? 0 : (ib >= g->res[k] ? g->res[k] - 1 : ib);
    }
}

static void build_grid(p_scene* ps) {
    p_grid* g = &ps->grid;
    memset(g, 0, sizeof *g);
    if (g_accel == 1 || ps->n_prims == 0 || (g_accel == 0 && ps->n_prims < 512)) return;
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = 0; i < ps->n_prims; ++i) {
        double a[3], b[3];
        prim_box(&ps->prims[i], a, b);
        for (int k = 0; k < 3; ++k) { lo[k] = fmin(lo[k], a[k]); hi[k] = fmax(hi[k], b[k]); }
    }
    double ext[3], vol = 1;
    for (int k = 0; k < 3; ++k) {
        double pad = 1e-6 * fmax(1.0, hi[k] - lo[k]);
        lo[k] -= pad; hi[k] += pad;
        ext[k] = hi[k] - lo[k];
        vol *= ext[k];
    }
    double s = cbrt(vol / (2.0 * ps->n_prims)); /* about two cells per primitive */
    size_t cells = 1;
    for (int k = 0; k < 3; ++k) {
        int r = (int)ceil(ext[k] / s);
        g->res[k] = r < 1 ? 1 : (r > 512 ? 512 : r);
        g->lo[k] = lo[k];
        g->cell[k] = ext[k] / g->res[k];
        g->inv_cell[k] = 1.0 / g->cell[k];
        cells *= (size_t)g->res[k];
    }
    g->start = calloc(cells + 1, sizeof(int));
    for (int pass = 0; pass < 2; ++pass) { /* count, then fill (primitive order is kept inside a cell) */
        for (int i = 0; i < ps->n_prims; ++i) {
            double a[3], b[3];
            int c0[3], c1[3];
            prim_box(&ps->prims[i], a, b);
            cell_range(g, a, b, c0, c1);
            for (int z = c0[2]; z <= c1[2]; ++z)
                for (int y = c0[1]; y <= c1[1]; ++y)
                    for (int x = c0[0]; x <= c1[0]; ++x) {
                        size_t c = ((size_t)z * g->res[1] + y) * g->res[0] + x;
                        if (pass == 0) g->start[c + 1]++;
                        else g->refs[g->start[c]++] = i;
                    }
        }
        if (pass == 0) {
            for (size_t c = 0; c < cells; ++c) g->start[c + 1] += g->start[c];
            g->refs = malloc(((size_t)g->start[cells] + 1) * sizeof(int));
        } else {
            for (size_t c = cells; c > 0; --c) g->start[c] = g->start[c - 1]; /* undo the fill cursor shift */
            g->start[0] = 0;
        }
    }
    g->on = 1;
}

/* G19_SHAPES_FIXED: the triangles the composite constructors MEANT to build (see the table in
 * 2019global_b200/csrc/fixed_shapes.cpp; reference include/entities.h:308-324, 379-406, 457-506, 579-590, 821-899).
 * Written against the reference's constructor arguments, in the same operation order as the product's generator so
 * that both round to the same float vertices. Returns the number of triangles (9 doubles each) or -1 when the
 * reference's own triangles are right. */
static int fixed_tris(const g19_entity_desc* d, double* out /* up to 200 x 9 */) {
    const double kPi = 3.14159265358979323846;
    const double px = d->p[0], py = d->p[1], pz = d->p[2];
    int n = 0;
#define TRI(ax, ay, az, bx, by, bz, cx, cy, cz)                                                  \
    do {                                                                                         \
        double* t_ = out + 9 * n++;                                                              \
        t_[0] = ax; t_[1] = ay; t_[2] = az; t_[3] = bx; t_[4] = by; t_[5] = bz; t_[6] = cx; t_[7] = cy; t_[8] = cz; \
    } while (0)
#define RECT(ax, ay, az, bx, by, bz, cx, cy, cz)                                                 \
    do {                                                                                         \
        TRI(ax, ay, az, bx, by, bz, cx, cy, cz);                                                 \
        TRI(ax, ay, az, bx, by, bz, ((ax) + (bx)) - (cx), ((ay) + (by)) - (cy), ((az) + (bz)) - (cz)); \
    } while (0)
    switch (d->kind) {
    case G19_EXP_RECTANGLE:
        RECT(px, py, pz, d->p[3], d->p[4], d->p[5], d->p[6], d->p[7], d->p[8]);
        return n;
    case G19_EXP_BOX: {
        const double x0 = px, y0 = py, z0 = pz, x1 = d->p[3], y1 = d->p[4], z1 = d->p[5];
        RECT(x0, y0, z0, x1, y1, z0, x1, y0, z0);
        RECT(x0, y0, z1, x1, y1, z1, x1, y0, z1);
        RECT(x0, y0, z0, x1, y0, z1, x1, y0, z0);
        RECT(x0, y1, z0, x1, y1, z1, x1, y1, z0);
        RECT(x0, y0, z0, x0, y1, z1, x0, y1, z0);
        RECT(x1, y0, z0, x1, y1, z1, x1, y1, z0);
        return n;
    }
    case G19_EXP_SPHERE: {
        const double r = (double)d->f[0];
        enum { STACKS = 10, SECTORS = 10 };
        double v[(STACKS + 1) * (SECTORS + 1)][3];
        int k = 0;
        for (int i = 0; i <= STACKS; ++i) {
            const double phi = kPi / 2 - (double)i * (kPi / STACKS);
            const double ring = r * cos(phi), z = r * sin(phi);
            for (int j = 0; j <= SECTORS; ++j, ++k) {
                const double theta = (double)j * (2 * kPi / SECTORS);
                v[k][0] = px + ring * cos(theta); v[k][1] = py + ring * sin(theta); v[k][2] = pz + z;
            }
        }
        for (int i = 0; i < STACKS; ++i) {
            int k1 = i * (SECTORS + 1), k2 = k1 + SECTORS + 1;
            for (int j = 0; j < SECTORS; ++j, ++k1, ++k2) {
                if (i != 0) TRI(v[k1][0], v[k1][1], v[k1][2], v[k2][0], v[k2][1], v[k2][2], v[k1 + 1][0], v[k1 + 1][1], v[k1 + 1][2]);
                if (i != STACKS - 1) TRI(v[k1 + 1][0], v[k1 + 1][1], v[k1 + 1][2], v[k2][0], v[k2][1], v[k2][2], v[k2 + 1][0], v[k2 + 1][1], v[k2 + 1][2]);
            }
        }
        return n;
    }
    case G19_EXP_QUAD: {
        const double hw = (double)d->f[0] / 2, hl = (double)d->f[1] / 2, a = (double)d->f[2];
        const double c = cos(a), s = sin(a);
        const double v0[3] = {px + hw * c, py + hl, pz + hw * s}, v1[3] = {px - hw * c, py + hl, pz - hw * s};
        const double v2[3] = {px + hw * c, py - hl, pz + hw * s}, v3[3] = {px - hw * c, py - hl, pz - hw * s};
        TRI(v1[0], v1[1], v1[2], v2[0], v2[1], v2[2], v0[0], v0[1], v0[2]);
        TRI(v1[0], v1[1], v1[2], v3[0], v3[1], v3[2], v2[0], v2[1], v2[2]);
        return n;
    }
    case G19_EXP_CONE: {
        const double h = (double)d->f[0], r = (double)d->f[1];
        v3 axis = V(d->p[3], d->p[4], d->p[5]);
        double l = sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
        if (l > 0) axis = vmul(axis, 1.0 / l);
        if (axis.x == 0 && axis.y == 0 && axis.z == 0) axis = V(0, 0, -1);
        const v3 centre = vadd(V(px, py, pz), vmul(axis, h));
        const v3 helper = fabs(axis.x) < 0.9 ? V(1, 0, 0) : V(0, 1, 0);
        v3 u = vcross(axis, helper);
        l = sqrt(u.x * u.x + u.y * u.y + u.z * u.z);
        if (l > 0) u = vmul(u, 1.0 / l);
        const v3 w = vcross(axis, u);
        enum { N = 50 };
        double rim[N + 1][3];
        for (int i = 0; i <= N; ++i) {
            const double ang = (double)i * (2 * kPi / N);
            const double cu = r * cos(ang), cw = r * sin(ang);
            rim[i][0] = centre.x + cu * u.x + cw * w.x; rim[i][1] = centre.y + cu * u.y + cw * w.y; rim[i][2] = centre.z + cu * u.z + cw * w.z;
        }
        for (int i = 0; i < N; ++i) {
            TRI(px, py, pz, rim[i][0], rim[i][1], rim[i][2], rim[i + 1][0], rim[i + 1][1], rim[i + 1][2]);
            TRI(centre.x, centre.y, centre.z, rim[i][0], rim[i][1], rim[i][2], rim[i + 1][0], rim[i + 1][1], rim[i + 1][2]);
        }
        return n;
    }
    default: return -1;
    }
#undef TRI
#undef RECT
}

void g19o_scene_set_shapes(void* h, int shapes) { ((o_scene*)h)->shapes = shapes; }

static p_scene* build_prims(const o_scene* s) {
    p_scene* ps = calloc(1, sizeof *ps);
    int cap = 0;
    for (int i = 0; i < s->n_ent; ++i) cap += (s->ent[i].ntri > 200 ? s->ent[i].ntri : 200) + 1;
    ps->prims = calloc((size_t)cap + 1, sizeof(p_prim));
    ps->lights = calloc((size_t)cap + 1, sizeof(p_light));
    for (int i = 0; i < s->n_ent; ++i) {
        const o_entity* e = &s->ent[i];
        if (!e->in_tree) continue;
        p_prim base;
        memset(&base, 0, sizeof base);
        base.bsdf = (e->desc.bsdf < 0 || e->desc.bsdf > 3) ? 0 : e->desc.bsdf;
        base.albedo = V(clamp01(e->desc.color[0]), clamp01(e->desc.color[1]), clamp01(e->desc.color[2]));
        base.emission = V(e->desc.emission[0], e->desc.emission[1], e->desc.emission[2]);
        base.ior = e->desc.ior > 0 ? e->desc.ior : 1.5;
        base.entity = i;
        if (e->kind == G19_IMP_SPHERE) {
            p_prim p = base;
            p.is_tri = 0;
            p.c = fround(e->pos);
            p.r = (double)e->radius;
            ps->prims[ps->n_prims++] = p;
            continue;
        }
        int from = (e->kind == G19_EXP_SPHERE) ? 1 : 0; /* the triangles REF mode tests */
        double fx[200 * 9];
        int ntri = e->ntri;
        const int nfixed = s->shapes == G19_SHAPES_FIXED ? fixed_tris(&e->desc, fx) : -1;
        if (nfixed >= 0) { from = 0; ntri = nfixed; }
        for (int t = from; t < ntri; ++t) {
            p_prim p = base;
            p.is_tri = 1;
            d3 q1, q2, q3;
            if (nfixed >= 0) {
                const double* f = fx + 9 * t;
                q1 = (d3){f[0], f[1], f[2]}; q2 = (d3){f[3], f[4], f[5]}; q3 = (d3){f[6], f[7], f[8]};
            } else {
                q1 = e->tris[t].p1; q2 = e->tris[t].p2; q3 = e->tris[t].p3;
            }
            v3 a = fround(q1), b = fround(q2), c = fround(q3);
            p.v0 = a;
            p.e1 = V((double)(float)(b.x - a.x), (double)(float)(b.y - a.y), (double)(float)(b.z - a.z));
            p.e2 = V((double)(float)(c.x - a.x), (double)(float)(c.y - a.y), (double)(float)(c.z - a.z));
            v3 n = vcross(V(q2.x - q1.x, q2.y - q1.y, q2.z - q1.z), V(q3.x - q1.x, q3.y - q1.y, q3.z - q1.z));
            double len = sqrt(vdot(n, n));
            if (len > 0) n = vmul(n, 1.0 / len);
            p.n = fround((d3){n.x, n.y, n.z});
            ps->prims[ps->n_prims++] = p;
            if (p.bsdf == G19_BSDF_EMITTER) {
                p_light l;
                l.v0 = p.v0; l.e1 = p.e1; l.e2 = p.e2; l.n = p.n; l.emission = p.emission;
                v3 cr = vcross(p.e1, p.e2);
                l.area = (double)(0.5f * sqrtf((float)vdot(cr, cr)));
                if (l.area > 0) ps->lights[ps->n_lights++] = l;
            }
        }
    }
    build_grid(ps);
    return ps;
}

static void free_prims(p_scene* ps) { free(ps->prims); free(ps->lights); free(ps->grid.start); free(ps->grid.refs); free(ps); }

/* Philox4x32-7 (Salmon et al. 2011: the 7-round variant is the fastest one that passes BigCrush; the 10-round one
 * used in round 1 spent 36 more instructions per path vertex on the device), key = (seed, "2019") */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t out[4]) {
    uint32_t k1 = 0x32303139u;
    for (int r = 0; r < 7; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static float u01f(uint32_t v) { return (float)(v >> 8) * (1.0f / 16777216.0f); }

static double hit_prim(const p_prim* p, v3 o, v3 d, double tmin, double tmax) {
    if (p->is_tri) {
        v3 pv = vcross(d, p->e2);
        double det = vdot(p->e1, pv);
        if (fabs(det) < 1.0e-20) return -1;
        double inv = 1.0 / det;
        v3 tv = vsub(o, p->v0);
        double u = vdot(tv, pv) * inv;
        if (u < 0 || u > 1) return -1;
        v3 qv = vcross(tv, p->e1);
        double v = vdot(d, qv) * inv;
        if (v < 0 || u + v > 1) return -1;
        double t = vdot(p->e2, qv) * inv;
        return (t > tmin && t < tmax) ? t : -1;
    }
    v3 oc = vsub(o, p->c);
    double b = vdot(oc, d);
    v3 l = vsub(oc, vmul(d, b));
    double disc = p->r * p->r - vdot(l, l);
    if (disc < 0) return -1;
    double sq = sqrt(disc), t0 = -b - sq, t1 = -b + sq;
    if (t0 > tmin && t0 < tmax) return t0;
    if (t1 > tmin && t1 < tmax) return t1;
    return -1;
}

/* 3-D DDA (Amanatides & Woo 1987) over the grid. visit(cell) is inlined in the two callers below. */
typedef struct {
    int c[3], step[3], out[3];
    double tmax[3], tdelta[3], t;
} dda;

static int dda_init(const p_grid* g, v3 o, v3 d, double tlimit, dda* w) {
    double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    double t0 = 0.0, t1 = tlimit;
    for (int k = 0; k < 3; ++k) { /* clip the ray to the grid box */
        double hi = g->lo[k] + g->cell[k] * g->res[k];
        if (dd[k] == 0.0) {
            if (oo[k] < g->lo[k] || oo[k] > hi) return 0;
        } else {
            double a = (g->lo[k] - oo[k]) / dd[k], b = (hi - oo[k]) / dd[k];
            if (a > b) { double q = a; a = b; b = q; }
            if (a > t0) t0 = a;
            if (b < t1) t1 = b;
        }
    }
    if (t0 > t1) return 0;
    w->t = t0;
    for (int k = 0; k < 3; ++k) {
        double pk = oo[k] + t0 * dd[k];
        int c = (int)floor((pk - g->lo[k]) * g->inv_cell[k]);
        c = c < 0 ? 0 : (c >= g->res[k] ? g->res[k] - 1 : c);
        w->c[k] = c;
        if (dd[k] > 0) {
            w->step[k] = 1; w->out[k] = g->res[k];
            w->tdelta[k] = g->cell[k] / dd[k];
            w->tmax[k] = (g->lo[k] + (c + 1) * g->cell[k] - oo[k]) / dd[k];
        } else if (dd[k] < 0) {
            w->step[k] = -1; w->out[k] = -1;
            w->tdelta[k] = -g->cell[k] / dd[k];
            w->tmax[k] = (g->lo[k] + c * g->cell[k] - oo[k]) / dd[k];
        } else {
            w->step[k] = 0; w->out[k] = -1;
            w->tdelta[k] = DBL_MAX; w->tmax[k] = DBL_MAX;
        }
    }
    return 1;
}
/* leaves the current cell; returns 0 when the ray leaves the grid. On return w->t = parameter of the new cell's entry. */
static int dda_next(dda* w) {
    int k = (w->tmax[0] <= w->tmax[1]) ? (w->tmax[0] <= w->tmax[2] ? 0 : 2) : (w->tmax[1] <= w->tmax[2] ? 1 : 2);
    w->t = w->tmax[k];
    w->c[k] += w->step[k];
    if (w->c[k] == w->out[k]) return 0;
    w->tmax[k] += w->tdelta[k];
    return 1;
}
static double dda_exit(const dda* w) { return fmin(w->tmax[0], fmin(w->tmax[1], w->tmax[2])); }

static int closest(const p_scene* s, v3 o, v3 d, double* t_out) {
    double best = DBL_MAX;
    int bi = -1;
    if (s->grid.on) {
        const p_grid* g = &s->grid;
        dda w;
        if (dda_init(g, o, d, DBL_MAX, &w)) {
            const double slack = 0.02 * fmin(g->cell[0], fmin(g->cell[1], g->cell[2])); /* lists are padded by 0.01 cell */
            do {
                size_t c = ((size_t)w.c[2] * g->res[1] + w.c[1]) * g->res[0] + w.c[0];
                for (int r = g->start[c]; r < g->start[c + 1]; ++r) {
                    int i = g->refs[r];
                    double t = hit_prim(&s->prims[i], o, d, 0.0, DBL_MAX);
                    /* the brute-force loop keeps the FIRST primitive among equal t: lowest index */
                    if (t >= 0 && (t < best || (t == best && i < bi))) { best = t; bi = i; }
                }
                if (best < dda_exit(&w) - slack) break; /* nothing in a later cell can be nearer */
            } while (dda_next(&w));
        }
        *t_out = best;
        return bi;
    }
    for (int i = 0; i < s->n_prims; ++i) {
        double t = hit_prim(&s->prims[i], o, d, 0.0, best);
        if (t >= 0) { best = t; bi = i; }
    }
    *t_out = best;
    return bi;
}
static int occluded(const p_scene* s, v3 o, v3 d, double tmax) {
    if (s->grid.on) {
        const p_grid* g = &s->grid;
        dda w;
        if (!dda_init(g, o, d, tmax, &w)) return 0;
        const double slack = 0.02 * fmin(g->cell[0], fmin(g->cell[1], g->cell[2]));
        do {
            if (w.t > tmax + slack) break;
            size_t c = ((size_t)w.c[2] * g->res[1] + w.c[1]) * g->res[0] + w.c[0];
            for (int r = g->start[c]; r < g->start[c + 1]; ++r)
                if (hit_prim(&s->prims[g->refs[r]], o, d, 0.0, tmax) >= 0) return 1;
        } while (dda_next(&w));
        return 0;
    }
    for (int i = 0; i < s->n_prims; ++i)
        if (hit_prim(&s->prims[i], o, d, 0.0, tmax) >= 0) return 1;
    return 0;
}

static void onb(v3 n, v3* t, v3* b) {
    double s = copysign(1.0, n.z), a = -1.0 / (s + n.z), bb = n.x * n.y * a;
    *t = V(1.0 + s * n.x * n.x * a, s * bb, -s * n.x);
    *b = V(bb, s + n.y * n.y * a, -n.y);
}

typedef struct {
    const p_scene* s;
    v3 cpos, up, left, top_left;
    int w, h, spp, max_depth, x0, y0, x1, y1, t, nthreads;
    uint32_t seed;
    float* radiance;
    uint64_t extend, shadow;
    char pad[128]; /* keep each thread's counters on their own cache lines */
} job;

static v3 trace_path(job* j, int x, int y, uint32_t sample) {
    const p_scene* s = j->s;
    uint32_t pixel = (uint32_t)y * (uint32_t)j->w + (uint32_t)x, r[4];
    philox(pixel, sample, 0, 0, j->seed, r);
    double fx = ((double)x + (double)u01f(r[0])) * 0.0002, fy = ((double)y + (double)u01f(r[1])) * 0.0002;
    v3 o = j->cpos;
    v3 d = vnorm(vsub(vsub(j->top_left, vmul(j->left, fx)), vmul(j->up, fy)));
    v3 T = V(1, 1, 1), L = V(0, 0, 0);
    int specular = 1;
    for (int b = 0; b < j->max_depth; ++b) {
        double t;
        j->extend++;
        int pi = closest(s, o, d, &t);
        if (pi < 0) break;
        const p_prim* pr = &s->prims[pi];
        if (pr->bsdf == G19_BSDF_EMITTER) {
            if (b == 0 || specular) L = vadd(L, vmulv(T, pr->emission));
            break;
        }
        v3 p = vadd(o, vmul(d, t));
        v3 ng = pr->is_tri ? pr->n : vmul(vsub(p, pr->c), 1.0 / pr->r);
        int entering = vdot(ng, d) < 0;
        v3 nf = entering ? ng : vmul(ng, -1);
        philox(pixel, sample, (uint32_t)b, 1, j->seed, r);
        v3 no, nd;
        if (pr->bsdf == G19_BSDF_DIFFUSE) {
            if (s->n_lights > 0) {
                float pick = u01f(r[0]) * (float)s->n_lights; /* float product, like the device */
                int li = (int)pick;
                if (li > s->n_lights - 1) li = s->n_lights - 1;
                double u1 = (double)(pick - (float)li), u2 = (double)u01f(r[1]);
                const p_light* lt = &s->lights[li];
                double su = sqrt(u1), b1 = su * (1.0 - u2), b2 = su * u2;
                v3 yl = vadd(vadd(lt->v0, vmul(lt->e1, b1)), vmul(lt->e2, b2));
                v3 wv = vsub(yl, p);
                double dist2 = vdot(wv, wv), dist = sqrt(dist2);
                wv = vmul(wv, 1.0 / dist);
                double cs = vdot(nf, wv), cl = fabs(vdot(lt->n, wv));
                if (cs > 0 && cl > 0 && dist > 2.0 * RAY_EPS) {
                    j->shadow++;
                    if (!occluded(s, vadd(p, vmul(nf, RAY_EPS)), wv, dist - 2.0 * RAY_EPS)) {
                        double g = cs * cl * lt->area / (dist2 * (1.0 / (double)s->n_lights)) * (1.0 / PI_D);
                        L = vadd(L, vmul(vmulv(vmulv(T, pr->albedo), lt->emission), g));
                    }
                }
            }
            double u3 = (double)u01f(r[2]), u4 = (double)u01f(r[3]);
            double rr = sqrt(u3), phi = 2.0 * PI_D * u4;
            v3 tx, ty;
            onb(nf, &tx, &ty);
            nd = vnorm(vadd(vadd(vmul(tx, rr * cos(phi)), vmul(ty, rr * sin(phi))), vmul(nf, sqrt(fmax(0.0, 1.0 - u3)))));
            no = vadd(p, vmul(nf, RAY_EPS));
            T = vmulv(T, pr->albedo);
            specular = 0;
        } else if (pr->bsdf == G19_BSDF_MIRROR) {
            nd = vnorm(vsub(d, vmul(nf, 2.0 * vdot(d, nf))));
            no = vadd(p, vmul(nf, RAY_EPS));
            T = vmulv(T, pr->albedo);
            specular = 1;
        } else {
            double etai = entering ? 1.0 : pr->ior, etat = entering ? pr->ior : 1.0, eta = etai / etat;
            double cosi = fmin(1.0, -vdot(d, nf));
            double sin2t = eta * eta * fmax(0.0, 1.0 - cosi * cosi), F = 1.0, cost = 0.0;
            if (sin2t < 1.0) {
                cost = sqrt(1.0 - sin2t);
                double rs = (etai * cosi - etat * cost) / (etai * cosi + etat * cost);
                double rp = (etai * cost - etat * cosi) / (etai * cost + etat * cosi);
                F = 0.5 * (rs * rs + rp * rp);
            }
            if ((double)u01f(r[0]) < F) {
                nd = vnorm(vadd(d, vmul(nf, 2.0 * cosi)));
                no = vadd(p, vmul(nf, RAY_EPS));
            } else {
                nd = vnorm(vadd(vmul(d, eta), vmul(nf, eta * cosi - cost)));
                no = vsub(p, vmul(nf, RAY_EPS));
            }
            T = vmulv(T, pr->albedo);
            specular = 1;
        }
        if (!(T.x > 0 || T.y > 0 || T.z > 0)) break;
        o = no;
        d = nd;
    }
    return L;
}

static void* rows(void* arg) {
    job* shared = arg;
    job local = *shared; /* thread-private copy: the segment counters are hot */
    job* j = &local;
    for (int y = j->y0 + j->t; y < j->y1; y += j->nthreads) {
        for (int x = j->x0; x < j->x1; ++x) {
            v3 sum = V(0, 0, 0);
            for (int sidx = 0; sidx < j->spp; ++sidx) sum = vadd(sum, trace_path(j, x, y, (uint32_t)sidx));
            size_t i = ((size_t)y * (size_t)j->w + (size_t)x) * 3;
            j->radiance[i] = (float)(sum.x / j->spp);
            j->radiance[i + 1] = (float)(sum.y / j->spp);
            j->radiance[i + 2] = (float)(sum.z / j->spp);
        }
    }
    shared->extend = local.extend;
    shared->shadow = local.shadow;
    return NULL;
}

static void camera_basis(const g19_camera* cam, int w, v3* cpos, v3* up, v3* left, v3* top_left);

/* Primary-hit AOV: the reference's own ray (integer pixel corner, no jitter, raytracer.h:41-43) against every
 * PATH primitive, nearest hit with t > 0. ids = entity (push order), points / normals row-major AoS triples;
 * the normal is turned towards the ray for triangles (entities.h:239-246) and is (p - c) / r for spheres
 * (entities.h:94); a miss is id -1, point DBL_MAX, normal 0 like ref_restate.c's g19o_trace. */
int g19o_path_primary(void* scene, const g19_camera* cam, int w, int h, int32_t* ids, double* points, double* normals) {
    const o_scene* os = scene;
    p_scene* ps = build_prims(os);
    v3 cpos, up, left, top_left;
    camera_basis(cam, w, &cpos, &up, &left, &top_left);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * (size_t)w + (size_t)x;
            double fx = (double)x * 0.0002, fy = (double)y * 0.0002;
            v3 d = vnorm(vsub(vsub(top_left, vmul(left, fx)), vmul(up, fy)));
            double t;
            int pi = closest(ps, cpos, d, &t);
            v3 p = V(DBL_MAX, DBL_MAX, DBL_MAX), n = V(0, 0, 0);
            int id = -1;
            if (pi >= 0) {
                const p_prim* pr = &ps->prims[pi];
                id = pr->entity;
                p = vadd(cpos, vmul(d, t));
                if (pr->is_tri) n = vdot(pr->n, d) > 0 ? vmul(pr->n, -1) : pr->n;
                else n = vmul(vsub(p, pr->c), 1.0 / pr->r);
            }
            if (ids) ids[i] = id;
            if (points) { points[3 * i] = p.x; points[3 * i + 1] = p.y; points[3 * i + 2] = p.z; }
            if (normals) { normals[3 * i] = n.x; normals[3 * i + 1] = n.y; normals[3 * i + 2] = n.z; }
        }
    free_prims(ps);
    return 0;
}

/* Renders the window [x0,x1) x [y0,y1) of a w x h frame; pixels outside are left
 * untouched. segs[0] = extend segments, segs[1] = shadow segments. */
int g19o_path_render(void* scene, const g19_camera* cam, int w, int h, int spp, int max_depth, uint32_t seed, int x0,
                     int y0, int x1, int y1, float* radiance, uint64_t* segs, int nthreads) {
    const o_scene* os = scene;
    p_scene* ps = build_prims(os);
    v3 cpos, up, left, top_left;
    camera_basis(cam, w, &cpos, &up, &left, &top_left);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    static job jobs[256];
    pthread_t th[256];
    for (int k = 0; k < nthreads; ++k) {
        job j = {ps, cpos, up, left, top_left, w, h, spp, max_depth, x0, y0, x1, y1, k, nthreads, seed, radiance, 0, 0};
        jobs[k] = j;
    }
    if (nthreads == 1) rows(&jobs[0]);
    else {
        for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, rows, &jobs[k]);
        for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    }
    if (segs) {
        segs[0] = segs[1] = 0;
        for (int k = 0; k < nthreads; ++k) { segs[0] += jobs[k].extend; segs[1] += jobs[k].shadow; }
    }
    free_prims(ps);
    return 0;
}

/* camera basis: same expressions as ref_restate.c g19o_trace (raytracer.h:26-30) */
static void camera_basis(const g19_camera* cam, int w, v3* cpos_out, v3* up_out, v3* left_out, v3* top_left_out) {
    v3 cpos = V(cam->pos[0], cam->pos[1], cam->pos[2]);
    v3 up = V(0, 0, 1.0);
    v3 fwd = vsub(V(cam->look_at[0], cam->look_at[1], cam->look_at[2]), cpos);
    {
        double tx = fwd.x * fwd.x, ty = fwd.y * fwd.y, tz = fwd.z * fwd.z;
        fwd = vmul(fwd, 1.0 / sqrt(tx + ty + tz));
    }
    v3 left = V(up.y * fwd.z - fwd.y * up.z, up.z * fwd.x - fwd.z * up.x, up.x * fwd.y - fwd.x * up.y);
    {
        double tx = left.x * left.x, ty = left.y * left.y, tz = left.z * left.z;
        left = vmul(left, 1.0 / sqrt(tx + ty + tz));
    }
    v3 t = vadd(cpos, V(cam->focal * fwd.x, cam->focal * fwd.y, cam->focal * fwd.z));
    t = vadd(t, vmul(vmul(vmul(left, (double)w), 0.5), 0.0002));
    t = vadd(t, vmul(vmul(vmul(up, (double)w), 0.5), 0.0002));
    *cpos_out = cpos;
    *up_out = up;
    *left_out = left;
    *top_left_out = vsub(t, cpos);
}
