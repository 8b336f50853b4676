/* ref_restate.c -- plain-C restatement of the reference's per-pixel loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_scene.h). This is the CPU oracle the
 * CUDA path is compared against. It restates, operation for operation and in
 * the reference's own float/double mix, what RayTracer::run executes:
 *   camera basis + pixel ray      reference include/raytracer.h:26-30,41-43
 *   Octree build                  include/octree.h:20-30,75-129; bbox.h:10-52
 *   Octree candidate collection   include/octree.h:132-155
 *   node ray/box test             include/entities.h:379-440 (ExpBox), 308-336
 *   ImpSphere / ImpTriangle       include/entities.h:53-96, 150-249
 *   composites                    include/entities.h:514-536,596-620,736-760,906-930
 *   front-object selection        include/raytracer.h:47-74  (LAST hit wins)
 *   getTextureCoord               include/entities.h:108-130,277-303,346-365,...
 *   Blinn-Phong + checker         include/material.h:48-106
 *   RGB888 store                  include/image.h:14-16
 *   GLM arithmetic order          3rd_party/glm/detail/func_geometric.inl:54-96,
 *                                 func_matrix.inl:90-108,272-294, type_mat3x3.inl:437-443
 *
 * PINNED (tests/test_oracle_ref.py): byte-equal to the compiled, unmodified
 * reference (oracle/_ref/libg19ref.so) on the config-1 image (sha256
 * 9ab01294...bba85), on hit ids / points / normals of every procedural scene,
 * and on randomised per-entity probes. Compile with -ffp-contract=off.
 *
 * Integer conversion of NaN / out-of-range doubles is UB in the reference; on
 * x86-64 it yields INT_MIN (cvttsd2si), which is what to_int() returns.
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_scene.h"

#define REF_PI 3.1415926535 /* entities.h:16 */

/* ---- GLM-order vector helpers ------------------------------------------- */
static d3 D3(double x, double y, double z) { d3 r = {x, y, z}; return r; }
static d3 add(d3 a, d3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
static d3 sub(d3 a, d3 b) { return D3(a.x - b.x, a.y - b.y, a.z - b.z); }
static d3 muls(d3 a, double s) { return D3(a.x * s, a.y * s, a.z * s); }
static d3 smul(double s, d3 a) { return D3(s * a.x, s * a.y, s * a.z); }
static d3 neg(d3 a) { return D3(-a.x, -a.y, -a.z); }
static double dot(d3 a, d3 b) { double tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
static d3 cross(d3 x, d3 y) { return D3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
static d3 normalize(d3 v) { return muls(v, 1.0 / sqrt(dot(v, v))); }
static double length3(d3 v) { return sqrt(dot(v, v)); }
static double dmin(double a, double b) { return (b < a) ? b : a; } /* std::min */
static double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max */
static int to_int(double v) { return (v != v || v >= 2147483648.0 || v <= -2147483649.0) ? INT_MIN : (int)v; }
static d3 f3_to_d3(float x, float y, float z) { return D3((double)x, (double)y, (double)z); }

/* ---- ImpTriangle --------------------------------------------------------- */
static void tri_init(o_tri* t, d3 p1, d3 p2, d3 p3) {
    t->p1 = p1; t->p2 = p2; t->p3 = p3;
    t->edge1 = sub(p2, p1);
    t->edge2 = sub(p3, p1);
    t->normal = normalize(cross(t->edge1, t->edge2));
    t->bmin = D3(dmin(dmin(p1.x, p2.x), p3.x), dmin(dmin(p1.y, p2.y), p3.y), dmin(dmin(p1.z, p2.z), p3.z));
    t->bmax = D3(dmax(dmax(p1.x, p2.x), p3.x), dmax(dmax(p1.y, p2.y), p3.y), dmax(dmax(p1.z, p2.z), p3.z) + 0.01);
    t->pos = smul(0.5, add(smul(0.5, add(p1, p2)), p3));
}

static void tri_bbox(const o_tri* t, d3* mn, d3* mx) {
    double x = t->bmax.x, y = t->bmax.y, z = t->bmax.z;
    if (t->bmax.x == t->bmin.x) x += 1e-5;
    if (t->bmax.y == t->bmin.y) y += 1e-5;
    if (t->bmax.z == t->bmin.z) z += 1e-5;
    *mn = t->bmin;
    *mx = D3(x, y, z);
}

/* entities.h:150-249 */
static int tri_intersect(const o_tri* t, d3 o, d3 dir, d3* ip, d3* in) {
    if (dot(t->normal, dir) == 0) return 0;
    /* A = transpose(mat3(edge1, edge2, -dir)) in FLOAT: A[c][r] */
    d3 nd = neg(dir);
    float m00 = (float)t->edge1.x, m01 = (float)t->edge2.x, m02 = (float)nd.x;
    float m10 = (float)t->edge1.y, m11 = (float)t->edge2.y, m12 = (float)nd.y;
    float m20 = (float)t->edge1.z, m21 = (float)t->edge2.z, m22 = (float)nd.z;
    float ood = 1.0f / (+m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02));
    float i20 = +(m10 * m21 - m20 * m11) * ood;
    float i21 = -(m00 * m21 - m20 * m01) * ood;
    float i22 = +(m00 * m11 - m10 * m01) * ood;
    d3 right = sub(o, t->pos);
    float vx = (float)right.x, vy = (float)right.y, vz = (float)right.z;
    float solz = i20 * vx + i21 * vy + i22 * vz;
    d3 point = add(o, smul((double)solz, dir));

    d3 d1 = normalize(cross(sub(t->p1, point), sub(t->p2, point)));
    d3 d2 = normalize(cross(sub(t->p2, point), sub(t->p3, point)));
    d3 d3_ = normalize(cross(sub(t->p3, point), sub(t->p1, point)));
    double eps = 1.0e-3;
    int hit = 0;
    if (length3(d1) < eps) hit = 1;
    else if (length3(d2) < eps) hit = 1;
    else if (length3(d3_) < eps) hit = 1;
    else {
        d3 a = sub(d1, d2), b = sub(d2, d3_);
        int cp1 = a.x * a.x + a.y * a.y + a.z * a.z < eps;
        int cp2 = b.x * b.x + b.y * b.y + b.z * b.z < eps;
        hit = cp1 && cp2;
    }
    if (!hit) return 0;
    *ip = point;
    *in = (dot(dir, t->normal) < 0) ? t->normal : neg(t->normal);
    return 1;
}

/* entities.h:53-96 */
static int sphere_intersect(const o_entity* e, d3 o, d3 dir, d3* ip, d3* in) {
    d3 np = sub(e->pos, o);
    float a1 = 1, a2 = 1, a3 = 1;
    if (dir.x != 0) { a2 = (float)(dir.y / dir.x); a3 = (float)(dir.z / dir.x); }
    else if (dir.y != 0) { a1 = (float)(dir.x / dir.y); a3 = (float)(dir.z / dir.y); }
    else if (dir.z != 0) { a2 = (float)(dir.y / dir.z); a1 = (float)(dir.x / dir.z); }
    else return 0;
    float r = e->radius;
    float a = (float)((double)a1 * (double)a1 + (double)a2 * (double)a2 + (double)a3 * (double)a3);
    float b = (float)(-2 * (np.x * a1 + np.y * a2 + np.z * a3));
    float c = (float)(np.x * np.x + np.y * np.y + np.z * np.z - (double)r * (double)r);
    float fac = 4 * a * c; /* float product */
    double disc = (double)b * (double)b - fac;
    if (disc < 0) return 0;
    float v1 = (float)((-b + sqrt(disc)) / (2 * a));
    float v2 = (float)((-b - sqrt(disc)) / (2 * a));
    float f1 = fabsf(v1), f2 = fabsf(v2);
    float base = (f2 < f1) ? f2 : f1;
    d3 p = f3_to_d3(base * a1, base * a2, base * a3);
    p = add(p, o);
    *ip = p;
    *in = normalize(sub(p, e->pos));
    return 1;
}

/* nearest-with-<= over a triangle list: entities.h:596-620 et al. */
static int tris_nearest(const o_tri* t, int from, int n, d3 o, d3 dir, d3* ip, d3* in) {
    int flag = 0;
    double best = DBL_MAX;
    d3 bp = D3(DBL_MAX, DBL_MAX, DBL_MAX), bn = D3(0, 0, 0);
    for (int i = from; i < n; ++i) {
        d3 p, nn;
        if (tri_intersect(&t[i], o, dir, &p, &nn)) {
            d3 tp = sub(p, o);
            double dsq = tp.x * tp.x + tp.y * tp.y + tp.z * tp.z;
            if (dsq <= best) { bp = p; bn = nn; best = dsq; }
            flag = 1;
        }
    }
    *ip = bp; *in = bn;
    return flag;
}

static int entity_intersect(const o_entity* e, d3 o, d3 dir, d3* ip, d3* in) {
    switch (e->kind) {
    case G19_IMP_SPHERE: return sphere_intersect(e, o, dir, ip, in);
    case G19_IMP_TRIANGLE: return tri_intersect(&e->tris[0], o, dir, ip, in);
    case G19_EXP_RECTANGLE: /* entities.h:326-336 */
        if (tri_intersect(&e->tris[0], o, dir, ip, in)) return 1;
        if (tri_intersect(&e->tris[1], o, dir, ip, in)) return 1;
        return 0;
    case G19_EXP_BOX: { /* entities.h:415-440: every hitting face overwrites */
        int has = 0;
        for (int f = 0; f < 6; ++f) {
            d3 p = D3(0, 0, 0), nn = D3(0, 0, 0);
            int h = tri_intersect(&e->tris[2 * f], o, dir, &p, &nn);
            if (!h) h = tri_intersect(&e->tris[2 * f + 1], o, dir, &p, &nn);
            if (h) { *ip = p; *in = nn; has = 1; }
        }
        return has;
    }
    case G19_EXP_SPHERE: return tris_nearest(e->tris, 1, e->ntri, o, dir, ip, in); /* i starts at 1, :520 */
    case G19_EXP_QUAD:
    case G19_EXP_CUBE:
    case G19_EXP_CONE: return tris_nearest(e->tris, 0, e->ntri, o, dir, ip, in);
    }
    return 0;
}

/* ---- node test: ExpBox(min,max).intersect boolean (entities.h:381-440) --- */
static void box_tris(d3 mn, d3 mx, o_tri out[12]) {
    d3 dlb = mn, drb = D3(mx.x, mn.y, mn.z), dlt = D3(mn.x, mx.y, mn.z), drt = D3(mx.x, mx.y, mn.z);
    d3 ulb = D3(mn.x, mn.y, mx.z), urb = D3(mx.x, mn.y, mx.z), ult = D3(mn.x, mx.y, mx.z), urt = mx;
    d3 f[6][3] = {{dlb, urb, ulb}, {dlb, ult, dlt}, {dlb, drt, dlt}, {urt, ulb, ult}, {urt, drb, drt}, {urt, dlt, drt}};
    for (int i = 0; i < 6; ++i) {
        d3 p1 = f[i][0], p2 = f[i][1], p3 = f[i][2];
        /* p4 = pos + (pos - p3) evaluated while Entity::pos is still {0,0,0}
         * (entities.h:319 runs before the ctor body :312) */
        d3 z = D3(0, 0, 0);
        d3 p4 = add(z, sub(z, p3));
        tri_init(&out[2 * i], p1, p2, p3);
        tri_init(&out[2 * i + 1], p1, p2, p4);
    }
}

static int node_box_hit(d3 mn, d3 mx, d3 o, d3 dir) {
    o_tri t[12];
    box_tris(mn, mx, t);
    int has = 0;
    for (int i = 0; i < 12; ++i) {
        d3 p, n;
        if (tri_intersect(&t[i], o, dir, &p, &n)) has = 1;
    }
    return has;
}

/* ---- BoundingBox (bbox.h:33-52) ------------------------------------------ */
static int bbox_overlap(d3 amin, d3 amax, d3 bmin, d3 bmax) {
    d3 p1 = smul(0.5, add(amin, amax));
    d3 p2 = smul(0.5, add(bmin, bmax));
    d3 dd = sub(p1, p2);
    int xo = fabs(dd.x) < (0.5 * (amax.x - amin.x) + 0.5 * (bmax.x - bmin.x));
    int yo = fabs(dd.y) < (0.5 * (amax.y - amin.y) + 0.5 * (bmax.y - bmin.y));
    int zo = fabs(dd.z) < (0.5 * (amax.z - amin.z) + 0.5 * (bmax.z - bmin.z));
    return xo && yo && zo;
}
static int all_le(d3 a, d3 b) { return a.x <= b.x && a.y <= b.y && a.z <= b.z; }

/* ---- entity construction -------------------------------------------------- */
static d3 P(const double* p) { return D3(p[0], p[1], p[2]); }

/* glm::mat3 (float, column-major m[c][r]) times a dvec3 narrowed to vec3 */
static d3 m3f_mul(const float m[3][3], d3 v) {
    float x = (float)v.x, y = (float)v.y, z = (float)v.z;
    float rx = m[0][0] * x + m[1][0] * y + m[2][0] * z;
    float ry = m[0][1] * x + m[1][1] * y + m[2][1] * z;
    float rz = m[0][2] * x + m[1][2] * y + m[2][2] * z;
    return f3_to_d3(rx, ry, rz);
}

static void set_bbox_f(o_entity* e, double x0, double y0, double z0, double x1, double y1, double z1) {
    /* BoundingBox(glm::vec3(...), glm::vec3(...)): doubles narrowed to float */
    e->bbmin = f3_to_d3((float)x0, (float)y0, (float)z0);
    e->bbmax = f3_to_d3((float)x1, (float)y1, (float)z1);
}

static int entity_build(o_entity* e, const g19_entity_desc* d) {
    memset(e, 0, sizeof *e);
    e->desc = *d;
    e->kind = d->kind;
    e->color = P(d->color);
    memcpy(e->f, d->f, sizeof e->f);
    const d3 z0 = D3(0, 0, 0); /* Entity::pos while in-class initialisers run */
    switch (d->kind) {
    case G19_IMP_SPHERE: {
        e->pos = P(d->p);
        e->radius = d->f[0];
        float r = e->radius;
        set_bbox_f(e, z0.x - r, z0.y - r, z0.z - r, z0.x + r, z0.y + r, z0.z + r); /* entities.h:98-99 */
        return 0;
    }
    case G19_IMP_TRIANGLE: {
        e->ntri = 1;
        e->tris = calloc(1, sizeof(o_tri));
        tri_init(&e->tris[0], P(d->p), P(d->p + 3), P(d->p + 6));
        e->pos = e->tris[0].pos;
        tri_bbox(&e->tris[0], &e->bbmin, &e->bbmax);
        return 0;
    }
    case G19_EXP_RECTANGLE: {
        d3 p1 = P(d->p), p2 = P(d->p + 3), p3 = P(d->p + 6);
        d3 p4 = add(z0, sub(z0, p3));
        e->ntri = 2;
        e->tris = calloc(2, sizeof(o_tri));
        tri_init(&e->tris[0], p1, p2, p3);
        tri_init(&e->tris[1], p1, p2, p4);
        e->p3 = p3; e->p4 = p4;
        e->pos = smul(0.5, add(p1, p2));
        e->bbmin = D3(dmin(p1.x, p2.x), dmin(p1.y, p2.y), dmin(p1.z, p2.z));
        e->bbmax = D3(dmax(p1.x, p2.x), dmax(p1.y, p2.y), dmax(p1.z, p2.z));
        return 0;
    }
    case G19_EXP_BOX: {
        e->ntri = 12;
        e->tris = calloc(12, sizeof(o_tri));
        box_tris(P(d->p), P(d->p + 3), e->tris);
        e->pos = z0;
        e->bbmin = P(d->p); e->bbmax = P(d->p + 3);
        return 0;
    }
    case G19_EXP_SPHERE: { /* entities.h:461-506 */
        d3 pos = P(d->p);
        float radius = d->f[0];
        e->pos = pos; e->radius = radius;
        const int sectornum = 10, stacknum = 10;
        float sectorStep = (float)(2 * REF_PI / sectornum);
        float stackStep = (float)(REF_PI / stacknum);
        d3 verts[11 * 11];
        int nv = 0;
        for (int i = 0; i <= stacknum; ++i) {
            float stackAngle = (float)(REF_PI / 2 - i * stackStep);
            float tmp = radius * cosf(stackAngle);
            float zf = (float)(radius * sinf(stackAngle) - pos.z);
            for (int j = 0; j <= sectornum; ++j) {
                float sectorAngle = j * sectorStep;
                float xf = (float)(tmp * cosf(sectorAngle) - pos.x);
                float yf = (float)(tmp * sinf(sectorAngle) - pos.y);
                verts[nv++] = f3_to_d3(xf, yf, zf);
            }
        }
        e->tris = calloc(2 * stacknum * sectornum, sizeof(o_tri));
        int nt = 0;
        for (int i = 0; i < stacknum; ++i) {
            int k1 = i * (sectornum + 1), k2 = k1 + sectornum + 1;
            for (int j = 0; j < sectornum; ++j, ++k1, ++k2) {
                if (i != 0) tri_init(&e->tris[nt++], verts[k1], verts[k2], verts[k1 + 1]);
                if (i != stacknum - 1) tri_init(&e->tris[nt++], verts[k1 + 1], verts[k2], verts[k2 + 1]);
            }
        }
        e->ntri = nt;
        set_bbox_f(e, z0.x - radius, z0.y - radius, z0.z - radius, z0.x + radius, z0.y + radius, z0.z + radius);
        return 0;
    }
    case G19_EXP_QUAD: { /* entities.h:581-590 */
        d3 pos = P(d->p);
        float width = d->f[0], length = d->f[1], alpha = d->f[2];
        e->pos = pos;
        float ca = cosf(alpha), sa = sinf(alpha); /* cos(float) resolves to the float overload */
        d3 v0 = D3((pos.x + width / 2) * ca, pos.y + length / 2, pos.z + (pos.x + width / 2) * sa);
        d3 v1 = D3((pos.x - width / 2) * ca, pos.y + length / 2, pos.z + (pos.x - width / 2) * sa);
        d3 v2 = D3((pos.x + width / 2) * ca, pos.y - length / 2, pos.z + (pos.x + width / 2) * sa);
        d3 v3 = D3((pos.x - width / 2) * ca, pos.y - length / 2, pos.z + pos.z + (pos.x - width / 2) * sa);
        e->ntri = 2;
        e->tris = calloc(2, sizeof(o_tri));
        tri_init(&e->tris[0], v1, v2, v0);
        tri_init(&e->tris[1], v1, v3, v2);
        /* keep vertices(0), vertices(1) for getTextureCoord */
        e->p3 = v0; e->p4 = v1;
        set_bbox_f(e, z0.x - width / 2, z0.y - length / 2, z0.z, z0.x + width / 2, z0.y + length / 2,
                   z0.z + (z0.x + width / 2) * sa); /* entities.h:623-624 */
        return 0;
    }
    case G19_EXP_CUBE: { /* entities.h:652-727 */
        d3 pos = P(d->p);
        float w = d->f[0], l = d->f[1], h = d->f[2];
        e->pos = pos;
        d3 v[8] = {D3(pos.x - w / 2, pos.y - l / 2, pos.z - h / 2), D3(pos.x - w / 2, pos.y - l / 2, pos.z + h / 2),
                   D3(pos.x + w / 2, pos.y - l / 2, pos.z - h / 2), D3(pos.x + w / 2, pos.y - l / 2, pos.z + h / 2),
                   D3(pos.x - w / 2, pos.y + l / 2, pos.z + h / 2), D3(pos.x - w / 2, pos.y + l / 2, pos.z - h / 2),
                   D3(pos.x + w / 2, pos.y + l / 2, pos.z - h / 2), D3(pos.x + w / 2, pos.y + l / 2, pos.z + h / 2)};
        static const int idx[12][3] = {{0, 1, 2}, {3, 1, 2}, {4, 5, 7}, {7, 5, 6}, {1, 0, 4}, {4, 0, 5},
                                       {3, 7, 2}, {7, 6, 2}, {1, 4, 3}, {3, 4, 7}, {0, 5, 2}, {2, 5, 6}};
        e->ntri = 12;
        e->tris = calloc(12, sizeof(o_tri));
        for (int i = 0; i < 12; ++i) tri_init(&e->tris[i], v[idx[i][0]], v[idx[i][1]], v[idx[i][2]]);
        e->p3 = v[0];
        set_bbox_f(e, z0.x - w / 2, z0.y - l / 2, z0.z - h / 2, z0.x + w / 2, z0.y + l / 2, z0.z + h / 2);
        return 0;
    }
    case G19_EXP_CONE: { /* entities.h:823-899 */
        d3 pos = P(d->p);
        float height = d->f[0], radius = d->f[1];
        e->pos = pos; e->radius = radius;
        /* the ctor overwrites its own PARAMETER: dir = normalize({-1,0,-10}) (:825) */
        d3 dir = normalize(D3(-1, 0, -10));
        float rx[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, ry[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        double x_sign = (dir.y < 0) ? 1.0 : -1.0;
        d3 x_dir = D3(0, dir.y, dir.z);
        if (!(x_dir.x == 0 && x_dir.y == 0 && x_dir.z == 0)) {
            double xa = x_sign * acos(dot(normalize(x_dir), D3(0, 0, -1)));
            float m[3][3] = {{1, 0, 0}, {0, (float)cos(xa), (float)-sin(xa)}, {0, (float)sin(xa), (float)cos(xa)}};
            memcpy(rx, m, sizeof m);
        }
        double y_sign = (dir.x > 0) ? 1.0 : -1.0;
        d3 y_dir = D3(dir.x, 0, -sqrt(dir.z * dir.z + dir.y * dir.y));
        if (!(y_dir.x == 0 && y_dir.y == 0 && y_dir.z == 0)) {
            double ya = y_sign * acos(dot(normalize(y_dir), D3(0, 0, -1)));
            float m[3][3] = {{(float)cos(ya), 0, (float)sin(ya)}, {0, 1, 0}, {(float)-sin(ya), 0, (float)cos(ya)}};
            memcpy(ry, m, sizeof m);
        }
        d3 verts[52];
        int nv = 0;
        verts[nv++] = pos;
        d3 loc = D3(0, 0, 0);
        double numSub = 50.0;
        for (int i = 0; i <= numSub; ++i) {
            float alpha = (float)(i * 360.0 / numSub);
            loc.x = pos.x + radius * cos(alpha * REF_PI / 180.0);
            loc.y = pos.y + radius * sin(alpha * REF_PI / 180.0);
            loc.z = pos.z - height;
            loc = sub(loc, pos);
            loc = m3f_mul(rx, loc);
            loc = m3f_mul(ry, loc);
            loc = add(loc, pos);
            verts[nv++] = loc;
        }
        e->ntri = 100;
        e->tris = calloc(100, sizeof(o_tri));
        int nt = 0;
        for (int i = 1; i < nv - 1; ++i) {
            tri_init(&e->tris[nt++], pos, verts[i], verts[i + 1]);
            d3 bc = add(pos, muls(normalize(dir), (double)height));
            tri_init(&e->tris[nt++], bc, verts[i], verts[i + 1]);
        }
        set_bbox_f(e, z0.x - radius, z0.y - radius, z0.z - height, z0.x + radius, z0.y + radius, z0.z);
        return 0;
    }
    }
    return -1;
}

/* ---- getTextureCoord (per type) ------------------------------------------- */
static double sq(double x) { return x * x; }
static double len3(d3 v) { return sqrt(sq(v.x) + sq(v.y) + sq(v.z)); }

static void texcoord(const o_entity* e, d3 ip, int* ox, int* oy) {
    int x = 0, y = 0;
    switch (e->kind) {
    case G19_IMP_SPHERE:
    case G19_EXP_SPHERE: { /* entities.h:108-130 / 549-571 */
        double radius = (double)e->radius;
        double ulv = 2.0 * REF_PI * radius / 320.0;
        d3 ti = sub(ip, e->pos);
        d3 up = D3(0, 0, radius);
        double cos_vert = dot(ti, up) / (radius * radius);
        double atu = acos(cos_vert);
        if (e->kind == G19_IMP_SPHERE) y = to_int((radius * atu) / ulv);
        else y = to_int((0.5 * REF_PI * radius - radius * atu) / ulv);
        double small_r = radius * sin(atu);
        d3 lm = D3(0, small_r, 0);
        double cos_hori = dot(D3(ti.x, ti.y, 0), lm) / (small_r * small_r);
        double ulh = 2.0 * REF_PI * small_r / 320.0;
        x = to_int(small_r * acos(cos_hori) / ulh);
        break;
    }
    case G19_IMP_TRIANGLE: { /* entities.h:277-303 */
        const o_tri* t = &e->tris[0];
        d3 p2p1 = sub(t->p2, t->p1), p3p1 = sub(t->p3, t->p1), p3p2 = sub(t->p3, t->p2), ip1 = sub(ip, t->p1);
        double p2p1_len = len3(p2p1), ip1_len = len3(ip1);
        double theta = acos(dot(p2p1, ip1) / (p2p1_len * ip1_len));
        double ix_len = ip1_len * sin(theta);
        d3 v = smul(0.5, add(p2p1, p3p1));
        double v_len = len3(v);
        d3 h = smul(0.5, add(neg(p3p2), neg(p3p1)));
        double h_len = len3(h);
        double ulv = v_len / 160.0, ulh = h_len / 160.0;
        y = to_int(ip1_len / ulh);
        x = to_int(ix_len / ulv);
        break;
    }
    case G19_EXP_RECTANGLE: { /* entities.h:346-365 */
        d3 p1 = e->tris[0].p1;
        d3 p3p1 = sub(e->p3, p1), p4p1 = sub(e->p4, p1);
        double width = len3(p4p1), length = len3(p3p1);
        double ulv = width / 64.0, ulh = length / 64.0;
        d3 ip1 = sub(ip, p1);
        double ip1_len = len3(ip1);
        double cos_theta = acos(dot(ip1, p3p1) / (length * ip1_len));
        x = to_int(ip1_len * sin(acos(cos_theta)) / ulh);
        y = to_int(ip1_len * cos_theta / ulv);
        break;
    }
    case G19_EXP_BOX: x = 0; y = 0; break;
    case G19_EXP_QUAD: { /* entities.h:630-641 */
        double width = (double)e->f[0], length = (double)e->f[1];
        double ulv = width / 160.0, ulh = length / 160.0;
        d3 right = sub(e->p3, e->p4); /* vertices(0) - vertices(1) */
        d3 ip1 = sub(ip, e->p4);
        double ip1_len = len3(ip1);
        double theta = acos(dot(ip1, right) / (width * ip1_len));
        y = to_int(ip1_len * sin(theta) / ulh);
        x = to_int(ip1_len * cos(theta) / ulv);
        break;
    }
    case G19_EXP_CUBE: { /* entities.h:769-811 */
        double width = (double)e->f[0], length = (double)e->f[1];
        double ulv = width / 160.0, ulh = length / 160.0;
        d3 right = D3(0, width, 0);
        d3 ipv = sub(ip, e->p3); /* vertices(0) */
        double ip_len = len3(ipv);
        double theta = acos(dot(ipv, right) / (width * ip_len));
        y = to_int(ip_len * sin(theta) / ulh);
        x = to_int(ip_len * cos(theta) / ulv);
        break;
    }
    case G19_EXP_CONE: { /* entities.h:942-961 */
        double radius = (double)e->f[1], height = (double)e->f[0];
        double ulh = sqrt(radius * radius + height * height) / 320.0;
        d3 ipos = sub(ip, e->pos);
        double y_len = len3(ipos);
        y = to_int(y_len / ulh);
        /* center = glm::vec3{pos.x,pos.y,intersect.z}: narrowed to float */
        d3 center = f3_to_d3((float)e->pos.x, (float)e->pos.y, (float)ip.z);
        double theta = atan(radius / height);
        double r_prime = y_len * sin(theta);
        d3 left = f3_to_d3(0.0f, (float)r_prime, 0.0f);
        d3 ic = sub(ip, center);
        double ulv = 2.0 * REF_PI * r_prime / 320.0;
        double alpha = acos(dot(ic, left) / (r_prime * r_prime));
        if (alpha > REF_PI / 4.0) alpha = acos(dot(ic, neg(left)) / (r_prime * r_prime));
        x = to_int(r_prime * alpha / ulv);
        break;
    }
    }
    *ox = x; *oy = y;
}

/* ---- Material::blinn_phong_texture (material.h:48-106) --------------------- */
/* Material::blinn_phong_texture, material.h:48-62. shader_parameters / specular_color / specular_power are the
 * entity's Material fields (material.h:25-29): the defaults of Material(color) unless the descriptor carries the
 * caller's (g19_entity_desc::material_set). */
static d3 shade(const o_entity* e, d3 dir, d3 light, d3 ip, d3 normal, int u, int v) {
    const g19_entity_desc* m = &e->desc;
    const d3 color = e->color;
    const d3 shader = m->material_set ? P(m->shader_parameters) : D3(0.1, 0.7, 1);
    const d3 spec = m->material_set ? P(m->specular_color) : D3(1, 1, 1);
    const double power = m->material_set ? m->specular_power : 5.0;
    int i = u % 32, j = v % 32;
    /* Negative u/v (ExpSphere's lower hemisphere, entities.h:558) make i/j
     * negative: the reference then indexes pattern[i][j] out of its row. Inside
     * the 32x32x3 block that is an ordinary flat-offset read, which is restated
     * here; a flat offset outside the block reads the reference's stack and is
     * UNDEFINED -- restated as the white cell and excluded from parity tests. */
    int flat = i * 32 + j;
    if (flat >= 0 && flat < 1024) { i = flat / 32; j = flat % 32; } else { i = 0; j = 0; }
    d3 tex;
    if ((i <= 16 && j <= 16) || (i > 16 && j > 16)) tex = D3(1, 1, 1);
    else tex = D3((double)(int)color.x, (double)(int)color.y, (double)(int)color.z);
    d3 tdc = muls(tex, 0.5);
    d3 la = muls(tex, shader.x);
    d3 ldir = normalize(sub(light, ip));
    d3 ld = muls(smul(dmax(0.0, dot(normal, ldir)), tdc), shader.y);
    d3 bis = normalize(add(normalize(neg(dir)), normalize(sub(light, ip))));
    d3 ls = muls(smul(pow(dmax(0.0, dot(normal, bis)), power), spec), shader.z);
    d3 out = add(add(la, ld), ls);
    return D3(dmin(out.x, 1.0), dmin(out.y, 1.0), dmin(out.z, 1.0));
}

static void quantise(d3 c, uint8_t* rgb) { /* image.h:14-16 + QColor validity */
    int r = to_int(255 * c.x), g = to_int(255 * c.y), b = to_int(255 * c.z);
    int ok = r >= 0 && r <= 255 && g >= 0 && g <= 255 && b >= 0 && b <= 255;
    rgb[0] = ok ? (uint8_t)r : 0; rgb[1] = ok ? (uint8_t)g : 0; rgb[2] = ok ? (uint8_t)b : 0;
}

/* ---- octree (octree.h) ----------------------------------------------------- */
static o_node* node_new(d3 mn, d3 mx) {
    o_node* n = calloc(1, sizeof *n);
    n->bmin = mn; n->bmax = mx;
    return n;
}
static void node_append(o_node* n, int e) {
    if (n->n_ent == n->cap_ent) {
        n->cap_ent = n->cap_ent ? 2 * n->cap_ent : 4;
        n->ent = realloc(n->ent, sizeof(int) * (size_t)n->cap_ent);
    }
    n->ent[n->n_ent++] = e;
}
static void node_free(o_node* n) {
    if (!n) return;
    for (int i = 0; i < 8; ++i) node_free(n->child[i]);
    free(n->ent);
    free(n);
}

static void node_partition(const o_scene* s, o_node* n) { /* octree.h:75-110 */
    if (n->child[0]) return;
    d3 mid = muls(add(n->bmin, n->bmax), 0.5);
    int all_in = 1;
    for (int i = 0; i < n->n_ent; ++i) {
        const o_entity* e = &s->ent[n->ent[i]];
        int bl = all_le(e->bbmin, mid), tr = all_le(mid, e->bbmax);
        all_in = all_in && bl && tr;
    }
    if (all_in) return;
    d3 a = n->bmin, b = n->bmax, m = mid;
    n->child[0] = node_new(a, m);
    n->child[1] = node_new(D3(a.x, m.y, a.z), D3(m.x, b.y, m.z));
    n->child[2] = node_new(D3(m.x, a.y, a.z), D3(b.x, m.y, m.z));
    n->child[3] = node_new(D3(m.x, m.y, a.z), D3(b.x, b.y, m.z));
    n->child[4] = node_new(m, b);
    n->child[5] = node_new(D3(a.x, m.y, m.z), D3(m.x, b.y, b.z));
    n->child[6] = node_new(D3(m.x, a.y, m.z), D3(b.x, m.y, b.z));
    n->child[7] = node_new(D3(a.x, a.y, m.z), D3(m.x, m.y, b.z));
}

static void node_push(const o_scene* s, o_node* n, int ei) { /* octree.h:115-129 */
    node_append(n, ei);
    node_partition(s, n);
    if (!n->child[0]) return;
    const o_entity* e = &s->ent[ei];
    for (int c = 0; c < 8; ++c) {
        o_node* ch = n->child[c];
        if (all_le(ch->bmin, e->bbmin) && all_le(e->bbmax, ch->bmax)) node_push(s, ch, ei);
        else if (bbox_overlap(ch->bmin, ch->bmax, e->bbmin, e->bbmax)) node_append(ch, ei);
    }
}

typedef struct { int* v; int n, cap; } ivec;
static void ivec_push(ivec* a, int x) {
    if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 64; a->v = realloc(a->v, sizeof(int) * (size_t)a->cap); }
    a->v[a->n++] = x;
}

static void node_candidates(const o_node* n, d3 o, d3 dir, ivec* out) { /* octree.h:132-155 */
    if (!n->child[0]) {
        for (int i = 0; i < n->n_ent; ++i) ivec_push(out, n->ent[i]);
        return;
    }
    for (int c = 0; c < 8; ++c) {
        const o_node* ch = n->child[c];
        if (ch->n_ent == 0) continue;
        if (node_box_hit(ch->bmin, ch->bmax, o, dir)) node_candidates(ch, o, dir, out);
    }
}

/* ---- C API (mirrors oracle/ref_harness/ref_driver.cpp one to one) ---------- */
void* g19o_scene_create(const double mn[3], const double mx[3]) {
    o_scene* s = calloc(1, sizeof *s);
    s->rmin = P(mn); s->rmax = P(mx);
    s->root = node_new(s->rmin, s->rmax);
    return s;
}

void g19o_scene_destroy(void* h) {
    o_scene* s = h;
    if (!s) return;
    node_free(s->root);
    for (int i = 0; i < s->n_ent; ++i) free(s->ent[i].tris);
    free(s->ent);
    free(s);
}

int g19o_scene_add(void* h, const g19_entity_desc* d) {
    o_scene* s = h;
    if (s->n_ent == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 16;
        s->ent = realloc(s->ent, sizeof(o_entity) * (size_t)s->cap);
    }
    if (entity_build(&s->ent[s->n_ent], d) != 0) return -1;
    int id = s->n_ent++;
    o_entity* e = &s->ent[id];
    e->in_tree = bbox_overlap(s->root->bmin, s->root->bmax, e->bbmin, e->bbmax);
    if (e->in_tree) node_push(s, s->root, id); /* octree.h:20-30 */
    return id;
}

int g19o_entity_count(void* h) { return ((o_scene*)h)->n_ent; }

int g19o_entity_bbox(void* h, int idx, double out[6]) {
    const o_entity* e = &((o_scene*)h)->ent[idx];
    out[0] = e->bbmin.x; out[1] = e->bbmin.y; out[2] = e->bbmin.z;
    out[3] = e->bbmax.x; out[4] = e->bbmax.y; out[5] = e->bbmax.z;
    return 0;
}

int g19o_entity_triangles(void* h, int idx, int kind, double* out, int max_tris) {
    const o_entity* e = &((o_scene*)h)->ent[idx];
    (void)kind;
    for (int i = 0; i < e->ntri && i < max_tris; ++i) {
        const o_tri* t = &e->tris[i];
        double v[9] = {t->p1.x, t->p1.y, t->p1.z, t->p2.x, t->p2.y, t->p2.z, t->p3.x, t->p3.y, t->p3.z};
        memcpy(out + 9 * i, v, sizeof v);
    }
    return e->ntri;
}

int g19o_triangle_derived(const double p[9], double out[12]) {
    o_tri t;
    tri_init(&t, P(p), P(p + 3), P(p + 6));
    double v[12] = {t.pos.x, t.pos.y, t.pos.z, t.edge1.x, t.edge1.y, t.edge1.z,
                    t.edge2.x, t.edge2.y, t.edge2.z, t.normal.x, t.normal.y, t.normal.z};
    memcpy(out, v, sizeof v);
    return 0;
}

int g19o_intersect(void* h, int idx, int n, const double* o, const double* d, int32_t* hit, double* points,
                   double* normals) {
    const o_entity* e = &((o_scene*)h)->ent[idx];
    for (int i = 0; i < n; ++i) {
        d3 dir = normalize(P(d + 3 * i)); /* Ray ctor, ray.h:6 */
        d3 p = D3(0, 0, 0), nn = D3(0, 0, 0);
        hit[i] = entity_intersect(e, P(o + 3 * i), dir, &p, &nn);
        points[3 * i] = p.x; points[3 * i + 1] = p.y; points[3 * i + 2] = p.z;
        normals[3 * i] = nn.x; normals[3 * i + 1] = nn.y; normals[3 * i + 2] = nn.z;
    }
    return 0;
}

int g19o_texcoord(void* h, int idx, int n, const double* points, int32_t* uv) {
    const o_entity* e = &((o_scene*)h)->ent[idx];
    for (int i = 0; i < n; ++i) {
        int x, y;
        texcoord(e, P(points + 3 * i), &x, &y);
        uv[2 * i] = x; uv[2 * i + 1] = y;
    }
    return 0;
}

int g19o_shade(void* h, int idx, const double o[3], const double d[3], const double light[3], const double point[3],
               const double normal[3], int u, int v, double rgb[3]) {
    const o_entity* e = &((o_scene*)h)->ent[idx];
    (void)o;
    d3 c = shade(e, normalize(P(d)), P(light), P(point), P(normal), u, v);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
    return 0;
}

int g19o_candidates(void* h, const double o[3], const double d[3], int32_t* out, int max_out) {
    const o_scene* s = h;
    ivec c = {0, 0, 0};
    node_candidates(s->root, P(o), normalize(P(d)), &c);
    for (int i = 0; i < c.n && i < max_out; ++i) out[i] = c.v[i];
    int n = c.n;
    free(c.v);
    return n;
}

typedef struct {
    const o_scene* s;
    d3 cpos, up, left, top_left, light;
    int w, y0, y1, t, nthreads;
    int32_t* ids; double* points; double* normals; uint8_t* rgb;
    uint64_t node_tests, prim_tests;
} band_arg;

static void* band_run(void* p) {
    band_arg* a = p;
    const o_scene* s = a->s;
    ivec cand = {0, 0, 0};
    for (int y = a->y0 + a->t; y < a->y1; y += a->nthreads) {
        for (int x = 0; x < a->w; ++x) {
            /* raytracer.h:41-43 */
            d3 direction = sub(sub(a->top_left, muls(muls(a->left, (double)x), 0.0002)), muls(muls(a->up, (double)y), 0.0002));
            d3 dir = normalize(direction);
            cand.n = 0;
            node_candidates(s->root, a->cpos, dir, &cand);
            d3 ip = D3(DBL_MAX, DBL_MAX, DBL_MAX), nn = D3(0, 0, 0);
            int front = -1;
            for (int i = 0; i < cand.n; ++i) { /* raytracer.h:53-74: every hit overwrites */
                d3 cp = D3(0, 0, 0), cn = D3(0, 0, 0);
                if (entity_intersect(&s->ent[cand.v[i]], a->cpos, dir, &cp, &cn)) { ip = cp; nn = cn; front = cand.v[i]; }
            }
            size_t i = (size_t)y * (size_t)a->w + (size_t)x;
            if (a->ids) a->ids[i] = front;
            if (a->points) { a->points[3 * i] = ip.x; a->points[3 * i + 1] = ip.y; a->points[3 * i + 2] = ip.z; }
            if (a->normals) { a->normals[3 * i] = nn.x; a->normals[3 * i + 1] = nn.y; a->normals[3 * i + 2] = nn.z; }
            if (a->rgb) {
                d3 c = D3(0, 0, 0);
                if (front >= 0) {
                    int u, v;
                    texcoord(&s->ent[front], ip, &u, &v);
                    c = shade(&s->ent[front], dir, a->light, ip, nn, u, v);
                }
                quantise(c, a->rgb + 3 * i);
            }
        }
    }
    free(cand.v);
    return NULL;
}

int g19o_trace(void* h, const g19_camera* cam, const double light[3], int w, int hgt, int y0, int y1, int32_t* ids,
               double* points, double* normals, uint8_t* rgb, int nthreads) {
    const o_scene* s = h;
    (void)hgt;
    d3 cpos = P(cam->pos);
    d3 up = D3(0, 0, 1.0);
    d3 forward = normalize(sub(P(cam->look_at), cpos)); /* camera.h:8-10 */
    d3 left = normalize(cross(up, forward));             /* raytracer.h:28 */
    /* raytracer.h:29-30 */
    d3 t = add(cpos, smul(cam->focal, forward));
    t = add(t, muls(muls(muls(left, (double)w), 0.5), 0.0002));
    t = add(t, muls(muls(muls(up, (double)w), 0.5), 0.0002));
    d3 top_left = sub(t, cpos);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    band_arg args[256];
    pthread_t th[256];
    for (int k = 0; k < nthreads; ++k) {
        band_arg a = {s, cpos, up, left, top_left, P(light), w, y0, y1, k, nthreads, ids, points, normals, rgb, 0, 0};
        args[k] = a;
    }
    if (nthreads == 1) { band_run(&args[0]); return 0; }
    for (int k = 0; k < nthreads; ++k) pthread_create(&th[k], NULL, band_run, &args[k]);
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
    return 0;
}

/* raytracer.h:76-82 on its own: getTextureCoord + blinn_phong_texture + setPixel for pixels whose front object
 * (ids, points, normals: row-major, AoS triples) was found elsewhere -- tests feed it the PATH oracle's primary
 * hits to check that PATH mode's depth-0 slice reduces to the reference's shade. */
int g19o_shade_pixels(void* h, const g19_camera* cam, const double light[3], int w, int hgt, const int32_t* ids,
                      const double* points, const double* normals, uint8_t* rgb) {
    const o_scene* s = h;
    d3 cpos = P(cam->pos);
    d3 up = D3(0, 0, 1.0);
    d3 forward = normalize(sub(P(cam->look_at), cpos));
    d3 left = normalize(cross(up, forward));
    d3 t = add(cpos, smul(cam->focal, forward));
    t = add(t, muls(muls(muls(left, (double)w), 0.5), 0.0002));
    t = add(t, muls(muls(muls(up, (double)w), 0.5), 0.0002));
    d3 top_left = sub(t, cpos);
    for (int y = 0; y < hgt; ++y)
        for (int x = 0; x < w; ++x) {
            size_t i = (size_t)y * (size_t)w + (size_t)x;
            d3 c = D3(0, 0, 0);
            int front = ids[i];
            if (front >= 0 && front < s->n_ent) {
                d3 dir = normalize(sub(sub(top_left, muls(muls(left, (double)x), 0.0002)), muls(muls(up, (double)y), 0.0002)));
                d3 ip = D3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
                d3 nn = D3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
                int u, v;
                texcoord(&s->ent[front], ip, &u, &v);
                c = shade(&s->ent[front], dir, P(light), ip, nn, u, v);
            }
            quantise(c, rgb + 3 * i);
        }
    return 0;
}

/* whole-frame convenience with the same signature as g19ref_render */
int g19o_render(void* h, const g19_camera* cam, const double light[3], int w, int hgt, uint8_t* rgb) {
    return g19o_trace(h, cam, light, w, hgt, 0, hgt, NULL, NULL, NULL, rgb, 1);
}
