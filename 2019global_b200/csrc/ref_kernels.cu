// ref_kernels.cu -- G19_MODE_REF: the reference's per-pixel loop on sm_100a.
//
// COMPILED WITH -fmad=false. Every double/float operation below is written in
// the order the reference evaluates it, so that hit booleans, hit points and
// normals come out bit-identical to the CPU program:
//   pixel ray              reference include/raytracer.h:41-43, include/ray.h:6
//   octree candidates      include/octree.h:132-155 (children 0..7, skip empty)
//   node ray/box test      include/entities.h:381-440 via 326-336 (12 line tests)
//   ImpSphere::intersect   include/entities.h:53-96
//   ImpTriangle::intersect include/entities.h:150-249
//   composites             include/entities.h:514-536, 596-620, 736-760, 906-930
//   front object           include/raytracer.h:53-74 (every hit overwrites)
//   getTextureCoord        include/entities.h:108-130 ... 942-961
//   blinn_phong_texture    include/material.h:48-106
//   Image::setPixel        include/image.h:14-16
//
// Design: one thread per pixel, 32x32 tiles so a warp is 32 consecutive x of
// one row (coherent rays, coalesced stores). The reference concatenates the
// leaf lists in DFS order and lets the LAST intersecting candidate win; the
// kernel walks the same tree in REVERSE (children 7..0, leaf entries last to
// first) and stops at the first hit -- the identical answer with early
// termination. The node test is the reference's own: the OR of twelve
// line/triangle tests against triangles derived from the child's min/max
// (including its p4 = -p3 construction bug), derived on the fly in registers
// rather than stored (168 B x 12 per node would dwarf the 64 B node record).
#include <algorithm>

#include "kernels.h"

#include <cfloat>
#include <climits>

namespace g19 {
namespace {

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 mk(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 ld3(const double* p) { return D3{p[0], p[1], p[2]}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator-(D3 a) { return D3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return D3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return D3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) {
    double tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz; // glm: tmp.x + tmp.y + tmp.z
}
__device__ __forceinline__ D3 cross3(D3 a, D3 b) {
    return D3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
__device__ __forceinline__ D3 unit(D3 v) { return v * (1.0 / sqrt(dot3(v, v))); } // glm::normalize
__device__ __forceinline__ double sqr(double v) { return v * v; }
__device__ __forceinline__ double len3(D3 v) { return sqrt(sqr(v.x) + sqr(v.y) + sqr(v.z)); }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
// (int)double as x86 cvttsd2si evaluates it (UB in C++; INT_MIN on the reference's platform)
__device__ __forceinline__ int ref_int(double v) {
    return (v != v || v >= 2147483648.0 || v <= -2147483649.0) ? INT_MIN : __double2int_rz(v);
}

// An ImpTriangle in registers.
struct Tri {
    D3 p1, p2, p3, pos, n;
    float e1x, e1y, e1z, e2x, e2y, e2z;
};

__device__ __forceinline__ Tri tri_load(const RefTriD* __restrict__ t) {
    Tri r;
    r.p1 = ld3(t->p1);
    r.p2 = ld3(t->p2);
    r.p3 = ld3(t->p3);
    r.pos = ld3(t->pos);
    r.n = ld3(t->normal);
    r.e1x = t->e1[0]; r.e1y = t->e1[1]; r.e1z = t->e1[2];
    r.e2x = t->e2[0]; r.e2y = t->e2[1]; r.e2z = t->e2[2];
    return r;
}

// glm::length(d) < 1e-3 without paying for the sqrt in the common case.
__device__ __forceinline__ bool shorter_than_eps(D3 d) {
    double q = dot3(d, d);
    return (q < 1.0e-5) && (sqrt(q) < 1.0e-3);
}

// The inside test of ImpTriangle::intersect (entities.h:182-237) decides with a LOOSE threshold: the three
// unit normals d_i = normalize(cross(p_i - p, p_j - p)) must pairwise agree to |d_i - d_j|^2 < 1e-3. For a point
// inside the triangle the d_i are equal (|d_i - d_j|^2 ~ 1e-30), outside two of them are opposite (~4); only
// within ~1e-3 of an edge does the value come anywhere near the threshold. Three FP64 square roots and three
// FP64 divisions per test (the normalisations) are therefore almost always spent on a foregone conclusion.
// inside_verdict() decides from the UN-normalised cross products when the conclusion is safe:
//     |d1 - d2|^2 = 2 - 2 cos,  cos = (c1 . c2) / sqrt(|c1|^2 |c2|^2);   < 1e-3  <=>  cos > 1 - 5e-4
// and returns "undecided" whenever cos is within 5e-7 of that bound (a million times the rounding error of the
// exact evaluation), or a cross product is degenerate / not finite -- then the caller runs the reference's own
// arithmetic. Every hit boolean is still the reference's, bit for bit (tests/test_ref_gpu.py probes 20 000 rays
// per entity class, and whole frames at 1080p against the compiled reference).
__device__ __forceinline__ int inside_verdict(D3 c1, D3 c2, D3 c3) { // 1 inside, 0 outside, -1 undecided
    const double n1 = dot3(c1, c1), n2 = dot3(c2, c2), n3 = dot3(c3, c3);
    if (!(n1 > 1.0e-200 && n2 > 1.0e-200 && n3 > 1.0e-200 && n1 < 1.0e200 && n2 < 1.0e200 && n3 < 1.0e200)) return -1;
    const double s12 = dot3(c1, c2), s23 = dot3(c2, c3);
    const double k = (1.0 - 5.0e-4) * (1.0 - 5.0e-4), k_hi = k + 1.0e-6, k_lo = k - 1.0e-6;
    const double q12 = s12 * s12, q23 = s23 * s23, m12 = n1 * n2, m23 = n2 * n3;
    const bool yes12 = s12 > 0 && q12 > k_hi * m12, no12 = s12 <= 0 || q12 < k_lo * m12;
    const bool yes23 = s23 > 0 && q23 > k_hi * m23, no23 = s23 <= 0 || q23 < k_lo * m23;
    if (no12 || no23) return 0;   // c1 && c2 is false as soon as one of them certainly is
    if (yes12 && yes23) return 1;
    return -1;
}

// the reference's own evaluation (dead `length < eps` branches kept: they fire only through NaNs)
__device__ __noinline__ bool inside_exact(D3 c1, D3 c2, D3 c3) {
    D3 d1 = unit(c1), d2 = unit(c2), d3 = unit(c3);
    if (shorter_than_eps(d1) || shorter_than_eps(d2) || shorter_than_eps(d3)) return true;
    D3 a = d1 - d2, b = d2 - d3;
    bool c1ok = sqr(a.x) + sqr(a.y) + sqr(a.z) < 1.0e-3;
    bool c2ok = sqr(b.x) + sqr(b.y) + sqr(b.z) < 1.0e-3;
    return c1ok && c2ok;
}

// where the ray's LINE meets the triangle's plane: the float solve of entities.h:154-166
__device__ __forceinline__ D3 plane_point(float e1x, float e1y, float e1z, float e2x, float e2y, float e2z, D3 pos, D3 o, D3 dir) {
    // A = transpose(mat3(edge1, edge2, -dir)) in float; only row 2 of inverse(A) is consumed
    float m02 = float(-dir.x), m12 = float(-dir.y), m22 = float(-dir.z);
    float m00 = e1x, m10 = e1y, m20 = e1z;
    float m01 = e2x, m11 = e2y, m21 = e2z;
    float det = +m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02);
    float ood = 1.0f / det;
    float i20 = +(m10 * m21 - m20 * m11) * ood;
    float i21 = -(m00 * m21 - m20 * m01) * ood;
    float i22 = +(m00 * m11 - m10 * m01) * ood;
    D3 rhs = o - pos;
    float vx = float(rhs.x), vy = float(rhs.y), vz = float(rhs.z);
    float solz = i20 * vx + i21 * vy + i22 * vz;
    return o + double(solz) * dir;
}

// ImpTriangle::intersect, entities.h:150-249. `facing` = normal as returned.
__device__ __forceinline__ bool tri_hit(const Tri& t, D3 o, D3 dir, D3& point, D3& facing) {
    double nd = dot3(t.n, dir);
    if (nd == 0) return false;
    D3 p = plane_point(t.e1x, t.e1y, t.e1z, t.e2x, t.e2y, t.e2z, t.pos, o, dir);
    D3 c1 = cross3(t.p1 - p, t.p2 - p), c2 = cross3(t.p2 - p, t.p3 - p), c3 = cross3(t.p3 - p, t.p1 - p);
    const int verdict = inside_verdict(c1, c2, c3);
    const bool hit = verdict < 0 ? inside_exact(c1, c2, c3) : verdict != 0;
    if (!hit) return false;
    point = p;
    facing = (nd < 0) ? t.n : -t.n; // dot(ray.dir, normal) < 0 -- same products, same sum
    return true;
}

// The same test for a face triangle of a node box (octree.h:141-146 builds an ExpBox per child per ray and keeps
// only the boolean): the members ImpTriangle derives at construction (entities.h:138-148) are computed here, and
// the normal is only normalised when its product with the direction is too close to zero to call.
__device__ __forceinline__ bool node_tri_hit(D3 p1, D3 p2, D3 p3, D3 o, D3 dir) {
    D3 e1 = p2 - p1, e2 = p3 - p1;
    D3 n = cross3(e1, e2);
    // reference: dot(normalize(n), dir) == 0 -> no hit. The normalisation scales the three products by one factor
    // (relative error 1e-16 each), so a sum that is clearly non-zero stays non-zero.
    const double tx = n.x * dir.x, ty = n.y * dir.y, tz = n.z * dir.z;
    const double nd_un = tx + ty + tz;
    if (!(fabs(nd_un) > 1.0e-9 * (fabs(tx) + fabs(ty) + fabs(tz)))) {
        if (dot3(unit(n), dir) == 0) return false;
    }
    D3 pos = 0.5 * (0.5 * (p1 + p2) + p3);
    D3 p = plane_point(float(e1.x), float(e1.y), float(e1.z), float(e2.x), float(e2.y), float(e2.z), pos, o, dir);
    D3 c1 = cross3(p1 - p, p2 - p), c2 = cross3(p2 - p, p3 - p), c3 = cross3(p3 - p, p1 - p);
    const int verdict = inside_verdict(c1, c2, c3);
    return verdict < 0 ? inside_exact(c1, c2, c3) : verdict != 0;
}

// ImpSphere::intersect, entities.h:53-96
__device__ __forceinline__ bool sphere_hit(D3 centre, float radius, D3 o, D3 dir, D3& point, D3& normal) {
    D3 np = centre - o;
    float a1 = 1, a2 = 1, a3 = 1;
    if (dir.x != 0) { a2 = float(dir.y / dir.x); a3 = float(dir.z / dir.x); }
    else if (dir.y != 0) { a1 = float(dir.x / dir.y); a3 = float(dir.z / dir.y); }
    else if (dir.z != 0) { a2 = float(dir.y / dir.z); a1 = float(dir.x / dir.z); }
    else return false;
    float a = float(double(a1) * double(a1) + double(a2) * double(a2) + double(a3) * double(a3));
    float b = float(-2 * (np.x * a1 + np.y * a2 + np.z * a3));
    float c = float(np.x * np.x + np.y * np.y + np.z * np.z - double(radius) * double(radius));
    float fac = 4 * a * c;
    double disc = double(b) * double(b) - fac;
    if (disc < 0) return false;
    float v1 = float((-b + sqrt(disc)) / (2 * a));
    float v2 = float((-b - sqrt(disc)) / (2 * a));
    float f1 = fabsf(v1), f2 = fabsf(v2);
    float base = (f2 < f1) ? f2 : f1;
    D3 p = mk(double(base * a1), double(base * a2), double(base * a3));
    p = p + o;
    point = p;
    normal = unit(p - centre);
    return true;
}

// min-distance-with-"<=" over a triangle range, entities.h:596-620
__device__ bool nearest_of(const RefTriD* __restrict__ tris, int from, int n, D3 o, D3 dir, D3& point, D3& normal,
                           unsigned& tests) {
    bool flag = false;
    double best = DBL_MAX;
    D3 bp = mk(DBL_MAX, DBL_MAX, DBL_MAX), bn = mk(0, 0, 0);
    for (int i = from; i < n; ++i) {
        Tri t = tri_load(tris + i);
        D3 p, nn;
        ++tests;
        if (tri_hit(t, o, dir, p, nn)) {
            D3 tp = p - o;
            double dsq = sqr(tp.x) + sqr(tp.y) + sqr(tp.z);
            if (dsq <= best) { bp = p; bn = nn; best = dsq; }
            flag = true;
        }
    }
    point = bp;
    normal = bn;
    return flag;
}

__device__ bool entity_hit(const RefSceneD& s, int ei, D3 o, D3 dir, D3& point, D3& normal, unsigned& tests) {
    const RefEntityD* __restrict__ e = s.entities + ei;
    const RefTriD* __restrict__ tris = s.tris + e->tri_offset;
    switch (e->combine) {
    case 0: ++tests; return sphere_hit(ld3(e->pos), e->radius, o, dir, point, normal);
    case 1: { ++tests; Tri t = tri_load(tris); return tri_hit(t, o, dir, point, normal); }
    case 2: { // ExpRectangle: t1, else t2
        Tri t = tri_load(tris);
        ++tests;
        if (tri_hit(t, o, dir, point, normal)) return true;
        t = tri_load(tris + 1);
        ++tests;
        return tri_hit(t, o, dir, point, normal);
    }
    case 3: { // ExpBox: the last hitting face overwrites
        bool has = false;
        for (int f = 0; f < 6; ++f) {
            D3 p, nn;
            Tri t = tri_load(tris + 2 * f);
            ++tests;
            bool h = tri_hit(t, o, dir, p, nn);
            if (!h) { t = tri_load(tris + 2 * f + 1); ++tests; h = tri_hit(t, o, dir, p, nn); }
            if (h) { point = p; normal = nn; has = true; }
        }
        return has;
    }
    default: return nearest_of(tris, e->first_tested, e->tri_count, o, dir, point, normal, tests);
    }
}

// ExpBox(min,max).intersect(ray) as a boolean (octree.h:141-146). The twelve
// triangles are those of entities.h:399-406 with ExpRectangle's p4 = 0+(0-p3).
// One of them: face f (0..5), `second` = the (p1, p2, p4) triangle of that face's ExpRectangle.
__device__ __forceinline__ bool node_face_tri_hit(const RefNodeD* __restrict__ nd, int f, bool second, D3 o, D3 dir) {
    const D3 mn = ld3(nd->mn), mx = ld3(nd->mx);
    // faces: (dlb,urb,ulb) (dlb,ult,dlt) (dlb,drt,dlt) (urt,ulb,ult) (urt,drb,drt) (urt,dlt,drt)
    D3 p1 = (f < 3) ? mn : mx, p2, p3;
    switch (f) {
    case 0: p2 = mk(mx.x, mn.y, mx.z); p3 = mk(mn.x, mn.y, mx.z); break;
    case 1: p2 = mk(mn.x, mx.y, mx.z); p3 = mk(mn.x, mx.y, mn.z); break;
    case 2: p2 = mk(mx.x, mx.y, mn.z); p3 = mk(mn.x, mx.y, mn.z); break;
    case 3: p2 = mk(mn.x, mn.y, mx.z); p3 = mk(mn.x, mx.y, mx.z); break;
    case 4: p2 = mk(mx.x, mn.y, mn.z); p3 = mk(mx.x, mx.y, mn.z); break;
    default: p2 = mk(mn.x, mx.y, mn.z); p3 = mk(mx.x, mx.y, mn.z); break;
    }
    if (second) {
        const D3 origin = mk(0, 0, 0);
        p3 = origin + (origin - p3); // p4
    }
    return node_tri_hit(p1, p2, p3, o, dir);
}

__device__ bool node_box_hit(const RefNodeD* __restrict__ nd, D3 o, D3 dir) {
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
        if (node_face_tri_hit(nd, f, false, o, dir)) return true;
        if (node_face_tri_hit(nd, f, true, o, dir)) return true;
    }
    return false;
}

__device__ __forceinline__ bool local_to_pixel(const TileMap& m, int lp, int& x, int& y) {
    int lt = lp / kTilePix, in = lp - lt * kTilePix;
    int t = lt * m.world + m.rank;
    int ty = t / m.tiles_x, tx = t - ty * m.tiles_x;
    x = tx * kTile + (in & (kTile - 1));
    y = ty * kTile + (in >> 5);
    return x < m.w && y < m.h;
}

__device__ __forceinline__ D3 pixel_dir(const RefCamera& c, int x, int y) { // raytracer.h:41 + ray.h:6
    D3 tl = ld3(c.top_left), left = ld3(c.left), up = ld3(c.up);
    D3 d = tl - left * double(x) * 0.0002 - up * double(y) * 0.0002;
    return unit(d);
}

constexpr int kMaxRefDepth = 40;

// Reverse DFS with early termination == "last hit of the forward DFS list".
__device__ int trace_front(const RefSceneD& s, D3 o, D3 dir, D3& point, D3& normal, unsigned& node_tests,
                           unsigned& prim_tests) {
    int stack_node[kMaxRefDepth];
    signed char stack_next[kMaxRefDepth];
    int level = 0;
    stack_node[0] = 0;
    stack_next[0] = 7;
    while (level >= 0) {
        const RefNodeD* __restrict__ nd = s.nodes + stack_node[level];
        int first = nd->first_child;
        if (first < 0) {
            const int32_t* __restrict__ list = s.ents + nd->ent_offset;
            for (int i = nd->ent_count - 1; i >= 0; --i) {
                int ei = list[i];
                if (entity_hit(s, ei, o, dir, point, normal, prim_tests)) return ei;
            }
            --level;
            continue;
        }
        int c = stack_next[level];
        bool pushed = false;
        for (; c >= 0; --c) {
            const RefNodeD* __restrict__ ch = s.nodes + first + c;
            if (ch->ent_count == 0) continue;
            ++node_tests;
            if (node_box_hit(ch, o, dir)) {
                stack_next[level] = (signed char)(c - 1);
                if (level + 1 < kMaxRefDepth) {
                    ++level;
                    stack_node[level] = first + c;
                    stack_next[level] = 7;
                    pushed = true;
                }
                break;
            }
        }
        if (!pushed) --level;
    }
    return -1;
}

__global__ void __launch_bounds__(128) ref_visibility_kernel(RefSceneD s, RefCamera cam, TileMap map, int lp0, int lp1,
                                                             int32_t* __restrict__ ids, double* __restrict__ points,
                                                             double* __restrict__ normals,
                                                             unsigned long long* __restrict__ counters) {
    int lp = lp0 + blockIdx.x * blockDim.x + threadIdx.x; // the band [lp0, lp1) of this rank's local pixels
    if (lp >= lp1) return;
    int x, y;
    unsigned node_tests = 0, prim_tests = 0;
    int id = -1;
    D3 point = mk(DBL_MAX, DBL_MAX, DBL_MAX), normal = mk(0, 0, 0);
    if (local_to_pixel(map, lp, x, y)) {
        D3 o = ld3(cam.pos);
        D3 dir = pixel_dir(cam, x, y);
        id = trace_front(s, o, dir, point, normal, node_tests, prim_tests);
        if (id < 0) { point = mk(DBL_MAX, DBL_MAX, DBL_MAX); normal = mk(0, 0, 0); }
    }
    ids[lp] = id;
    if (points) {
        // SoA planes: coalesced 8-byte stores
        points[lp] = point.x; points[map.n_local_pix + lp] = point.y; points[2 * (size_t)map.n_local_pix + lp] = point.z;
        normals[lp] = normal.x; normals[map.n_local_pix + lp] = normal.y; normals[2 * (size_t)map.n_local_pix + lp] = normal.z;
    }
    if (counters) {
        for (int off = 16; off > 0; off >>= 1) {
            node_tests += __shfl_xor_sync(0xffffffffu, node_tests, off);
            prim_tests += __shfl_xor_sync(0xffffffffu, prim_tests, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 0, (unsigned long long)node_tests);
            atomicAdd(counters + 1, (unsigned long long)prim_tests);
        }
    }
}

// Three face-triangle tests of a node's children at once, stage by stage over small arrays: after unrolling, the three
// FP64 chains are independent instruction streams inside the same basic blocks, which is what lets the scheduler
// interleave them (three calls of node_face_tri_hit stay three serial chains: each one is its own branchy region).
// Test t = pass * 32 + lane: child t / 12, face (t % 12) / 2, second triangle of the face's ExpRectangle when odd.
// Same arithmetic per test as node_face_tri_hit / node_tri_hit. Returns the children (bits) this lane found hit.
__device__ __forceinline__ unsigned node_children_hits3(const RefNodeD* __restrict__ children, D3 o, D3 dir, unsigned lane,
                                                        unsigned& node_tests) {
    D3 p1[3], p2[3], p3[3];
    bool live[3];
    unsigned bit[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int t = k * 32 + int(lane);
        const int c = t / 12, tri = t - 12 * c, f = tri >> 1;
        const RefNodeD* __restrict__ ch = children + c;
        live[k] = ch->ent_count != 0;
        bit[k] = 1u << c;
        if (live[k] && tri == 0) ++node_tests;
        const D3 mn = ld3(ch->mn), mx = ld3(ch->mx);
        // corner codes per face (bit 0 / 1 / 2: take max.x / max.y / max.z), entities.h:399-406:
        // faces (dlb,urb,ulb) (dlb,ult,dlt) (dlb,drt,dlt) (urt,ulb,ult) (urt,drb,drt) (urt,dlt,drt)
        const unsigned c2 = (0x118f5u >> (3 * f)) & 7u; // p2: 5, 6, 3, 4, 1, 2
        const unsigned c3 = (0x1bc94u >> (3 * f)) & 7u; // p3: 4, 2, 2, 6, 3, 3
        p1[k] = (f < 3) ? mn : mx;
        p2[k] = mk((c2 & 1u) ? mx.x : mn.x, (c2 & 2u) ? mx.y : mn.y, (c2 & 4u) ? mx.z : mn.z);
        const D3 q3 = mk((c3 & 1u) ? mx.x : mn.x, (c3 & 2u) ? mx.y : mn.y, (c3 & 4u) ? mx.z : mn.z);
        const D3 origin = mk(0, 0, 0);
        const D3 p4 = origin + (origin - q3); // ExpRectangle::p4 with pos still at the origin (entities.h:319)
        p3[k] = (tri & 1) ? p4 : q3;
    }
    D3 e1[3], e2[3];
    bool dead[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        e1[k] = p2[k] - p1[k];
        e2[k] = p3[k] - p1[k];
        const D3 n = cross3(e1[k], e2[k]);
        const double tx = n.x * dir.x, ty = n.y * dir.y, tz = n.z * dir.z;
        const double nd_un = tx + ty + tz;
        dead[k] = false;
        if (!(fabs(nd_un) > 1.0e-9 * (fabs(tx) + fabs(ty) + fabs(tz)))) dead[k] = dot3(unit(n), dir) == 0; // rare: see node_tri_hit
    }
    D3 p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const D3 pos = 0.5 * (0.5 * (p1[k] + p2[k]) + p3[k]);
        p[k] = plane_point(float(e1[k].x), float(e1[k].y), float(e1[k].z), float(e2[k].x), float(e2[k].y), float(e2[k].z), pos, o, dir);
    }
    unsigned found = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const D3 c1 = cross3(p1[k] - p[k], p2[k] - p[k]), c2 = cross3(p2[k] - p[k], p3[k] - p[k]), c3 = cross3(p3[k] - p[k], p1[k] - p[k]);
        const int verdict = inside_verdict(c1, c2, c3);
        const bool in = verdict < 0 ? inside_exact(c1, c2, c3) : verdict != 0;
        if (live[k] && !dead[k] && in) found |= bit[k];
    }
    return found;
}

// ---- one WARP per ray -----------------------------------------------------------------------------
// What the capture of ref_visibility_kernel on the 1 002 528-entity heightfield showed (profiles/r02a_ref_visibility_ncu.json):
// 2.44 s for 518 400 rays, but only 5.6 G warp instructions at 1.54 of 32 lanes, warps active 3.6 % of the time -- the
// AVERAGE ray costs ~40 triangle tests, a handful of rays walk thousands of nodes of the reference's entity-losing
// tree with twelve FP64 triangle tests per child, and the frame waits for the longest serial chain. The work of one
// ray parallelises perfectly: the 8 x 12 face triangles of a node's children are independent tests (octree.h:141-146
// only keeps the boolean) and so are the entities of a leaf list (the LAST hit wins, raytracer.h:58: the highest list
// position among the hits). So a warp takes one ray: 96 node-triangle tests in three passes of 32 lanes, leaf lists
// 32 entities at a time from the back. Same device functions, same visiting order (children 7 -> 0, stack of
// unvisited hit children), so the ids, points and normals are the thread-per-ray kernel's, bit for bit.
// start / level_cap: the walk covers the subtree under node `start`, whose depth in the tree is kMaxRefDepth - level_cap
// (0 / kMaxRefDepth: the whole tree). budget > 0: give up after that many node expansions and return kWalkOverBudget
// (the ray goes to the heavy list). poll != nullptr: another warp of the same ray may publish a hit in a subtree that
// comes EARLIER in the reversed order (a lower rank than `rank`): this walk is moot then and returns kWalkMoot.
constexpr int kWalkOverBudget = -2, kWalkMoot = -3;
__device__ int trace_front_warp(const RefSceneD& s, D3 o, D3 dir, D3& point, D3& normal, int* st_node, int* st_mask,
                                unsigned& node_tests, unsigned& prim_tests, int start = 0, int level_cap = kMaxRefDepth,
                                int budget = 0, const unsigned* poll = nullptr, unsigned rank = 0) {
    const unsigned lane = threadIdx.x & 31u;
    int level = 0;
    int rounds = 0;
    if (lane == 0) { st_node[0] = start; st_mask[0] = -1; }
    __syncwarp();
    while (level >= 0) {
        const RefNodeD* __restrict__ nd = s.nodes + st_node[level];
        const int first = nd->first_child;
        if (first < 0) { // leaf: its entity list from the back, 32 at a time; the first hit met is the reference's last
            const int n = nd->ent_count;
            const int32_t* __restrict__ list = s.ents + nd->ent_offset;
            for (int base = 0; base < n; base += 32) {
                const int k = n - 1 - (base + int(lane));
                bool h = false;
                D3 p = mk(0, 0, 0), nn = mk(0, 0, 0);
                int ei = -1;
                if (k >= 0) {
                    ei = list[k];
                    h = entity_hit(s, ei, o, dir, p, nn, prim_tests);
                }
                const unsigned m = __ballot_sync(0xffffffffu, h);
                if (m) {
                    const int src = __ffs(int(m)) - 1;
                    point = mk(__shfl_sync(0xffffffffu, p.x, src), __shfl_sync(0xffffffffu, p.y, src), __shfl_sync(0xffffffffu, p.z, src));
                    normal = mk(__shfl_sync(0xffffffffu, nn.x, src), __shfl_sync(0xffffffffu, nn.y, src), __shfl_sync(0xffffffffu, nn.z, src));
                    return __shfl_sync(0xffffffffu, ei, src);
                }
            }
            --level;
            continue;
        }
        int mask = st_mask[level];
        __syncwarp();
        if (mask < 0) { // the box tests of this node's non-empty children: 8 x 12 triangle tests over the lanes
            // three tests per lane, evaluated as three interleaved FP64 chains: one warp walks a heavy ray alone (the
            // most expensive ray of the 1 M-entity frame tests 318 932 child boxes, profiles/r02e), and nothing else
            // hides the latency of ~200 dependent double-precision operations per test
            if (budget > 0 && ++rounds > budget) return kWalkOverBudget;
            if (poll) { // warp-uniform: every lane reads the same word
                if (*reinterpret_cast<const volatile unsigned*>(poll) < rank) return kWalkMoot;
            }
            const unsigned mine = node_children_hits3(s.nodes + first, o, dir, lane, node_tests);
            mask = int(__reduce_or_sync(0xffffffffu, mine));
        }
        if (mask == 0) {
            --level;
            continue;
        }
        const int c = 31 - __clz(mask); // children 7 -> 0: the reverse of the reference's DFS
        __syncwarp();
        if (lane == 0) {
            st_mask[level] = mask & ~(1 << c);
            if (level + 1 < level_cap) {
                st_node[level + 1] = first + c;
                st_mask[level + 1] = -1;
            }
        }
        __syncwarp();
        if (level + 1 < level_cap) ++level;
    }
    return -1;
}

constexpr int kRefWarps = 4; // warps per CTA of the warp-per-ray kernel
__global__ void __launch_bounds__(kRefWarps * 32) ref_visibility_warp_kernel(RefSceneD s, RefCamera cam, TileMap map, int lp0, int lp1,
                                                                            int32_t* __restrict__ ids, double* __restrict__ points,
                                                                            double* __restrict__ normals,
                                                                            unsigned long long* __restrict__ counters,
                                                                            unsigned* __restrict__ next, RefHeavyD heavy) {
    __shared__ int st_node[kRefWarps][kMaxRefDepth];
    __shared__ int st_mask[kRefWarps][kMaxRefDepth];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned node_tests = 0, prim_tests = 0;
    // rays differ in cost by four orders of magnitude (most stop at the first leaf, a few walk thousands of nodes): a
    // warp takes its next pixel from a counter when it is done, so the expensive ones spread over the machine
    while (true) {
        unsigned take = 0;
        if (lane == 0) take = atomicAdd(next, 1u);
        take = __shfl_sync(0xffffffffu, take, 0);
        if (take >= unsigned(lp1 - lp0)) break;
        const int lp = lp0 + int(take);
        int x, y;
        int id = -1;
        D3 point = mk(DBL_MAX, DBL_MAX, DBL_MAX), normal = mk(0, 0, 0);
        if (local_to_pixel(map, lp, x, y)) {
            D3 o = ld3(cam.pos);
            D3 dir = pixel_dir(cam, x, y);
            const unsigned before = node_tests;
            id = trace_front_warp(s, o, dir, point, normal, st_node[wib], st_mask[wib], node_tests, prim_tests, 0, kMaxRefDepth,
                                  heavy.budget);
            if (id == kWalkOverBudget) {
                // a heavy ray: ref_heavy_kernel spreads it over many warps (it starts over: the rounds spent here are
                // the price of finding out). No room left in the list: walk it to the end right here.
                unsigned at = 0;
                if (lane == 0) at = atomicAdd(heavy.count, 1u);
                at = __shfl_sync(0xffffffffu, at, 0);
                if (at < unsigned(heavy.cap)) {
                    if (lane == 0) heavy.lp[at] = lp;
                    id = -1;
                } else {
                    id = trace_front_warp(s, o, dir, point, normal, st_node[wib], st_mask[wib], node_tests, prim_tests);
                }
            }
            if (id < 0) { point = mk(DBL_MAX, DBL_MAX, DBL_MAX); normal = mk(0, 0, 0); } // (a heavy ray reads "no hit" until its hit is published)
            if (counters) { // profile: the most expensive ray of the launch, in child-box tests (lanes count disjoint children)
                unsigned mine = node_tests - before;
                for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
                if (lane == 0) atomicMax(counters + 3, (unsigned long long)mine);
            }
        }
        __syncwarp();
        if (lane == 0) {
            ids[lp] = id;
            if (points) {
                points[lp] = point.x; points[map.n_local_pix + lp] = point.y; points[2 * (size_t)map.n_local_pix + lp] = point.z;
                normals[lp] = normal.x; normals[map.n_local_pix + lp] = normal.y; normals[2 * (size_t)map.n_local_pix + lp] = normal.z;
            }
        }
    }
    if (counters) {
        for (int off = 16; off > 0; off >>= 1) {
            node_tests += __shfl_xor_sync(0xffffffffu, node_tests, off);
            prim_tests += __shfl_xor_sync(0xffffffffu, prim_tests, off);
        }
        if (lane == 0) {
            atomicAdd(counters + 0, (unsigned long long)node_tests);
            atomicAdd(counters + 1, (unsigned long long)prim_tests);
        }
    }
}

// ---- heavy rays: ONE ray over many warps ------------------------------------------------------------
// Rays differ in cost by five orders of magnitude on the 1 M-entity heightfield: the average ray expands 1.25 nodes,
// the heaviest 39 866 (318 932 child-box tests, ~10 us of dependent FP64 arithmetic each), and with one warp per ray a
// 1080p frame waited 380 ms for that single walk (profiles/r02_tuning_log.md). The reversed walk is a depth-first
// search whose answer is "the first hit in visiting order", so it splits by subtree: task (ray, code) walks the subtree
// whose path from the root is the four child choices in `code` -- digit 0 is child 7, the child visited first -- and the
// ray's answer is the hit of the LOWEST code that has one. Tasks are handed out in code order from a counter; a hit is
// published under a per-ray lock together with its rank, and every walk of a higher rank gives up as soon as it sees
// one (trace_front_warp's poll), which is the early termination of the serial walk. The child-box masks of the four
// levels above the subtrees are computed once per ray by whichever task gets there first and kept in a small table
// (0 = not computed yet; two warps racing compute the same value). Same device functions as the serial walk, same
// visiting order inside every subtree: ids, points and normals are bit for bit what one warp would have found.
__global__ void __launch_bounds__(kRefWarps * 32) ref_heavy_kernel(RefSceneD s, RefCamera cam, TileMap map, int32_t* __restrict__ ids,
                                                                  double* __restrict__ points, double* __restrict__ normals,
                                                                  unsigned long long* __restrict__ counters, RefHeavyD heavy) {
    __shared__ int st_node[kRefWarps][kMaxRefDepth];
    __shared__ int st_mask[kRefWarps][kMaxRefDepth];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_heavy = min(*heavy.count, unsigned(heavy.cap));
    if (n_heavy == 0) return;
    unsigned node_tests = 0, prim_tests = 0;
    const D3 o = ld3(cam.pos);
    while (true) {
        unsigned take = 0;
        if (lane == 0) take = atomicAdd(heavy.count + 1, 1u);
        take = __shfl_sync(0xffffffffu, take, 0);
        const unsigned h = take >> (3 * kHeavyLevel);
        if (h >= n_heavy) break;
        const unsigned code = take & unsigned(kHeavyTasks - 1);
        if (*reinterpret_cast<const volatile unsigned*>(heavy.best + h) < code) continue; // a hit earlier in the order exists already
        const int lp = heavy.lp[h];
        int x, y;
        local_to_pixel(map, lp, x, y);
        const D3 dir = pixel_dir(cam, x, y);
        unsigned short* cache = heavy.masks + size_t(h) * kHeavyCache;
        int node = 0, at = 0, level_base = 0, depth = 0;
        bool walk = true, leaf_here = false;
        for (int l = 0; l < kHeavyLevel; ++l) {
            const RefNodeD* __restrict__ nd = s.nodes + node;
            const int first = nd->first_child;
            const unsigned below = code & ((1u << (3 * (kHeavyLevel - l))) - 1u);
            if (first < 0) { // a leaf above the split level: the first task under it takes it, the others have nothing to do
                leaf_here = true;
                walk = below == 0u;
                break;
            }
            const unsigned digit = (code >> (3 * (kHeavyLevel - 1 - l))) & 7u;
            const int c = 7 - int(digit);
            unsigned m = reinterpret_cast<const volatile unsigned short*>(cache)[level_base + at];
            if (!(m & 0x100u)) {
                const unsigned mine = node_children_hits3(s.nodes + first, o, dir, unsigned(lane), node_tests);
                m = 0x100u | __reduce_or_sync(0xffffffffu, mine);
                if (lane == 0) reinterpret_cast<volatile unsigned short*>(cache)[level_base + at] = (unsigned short)m;
            }
            if (!((m >> c) & 1u)) { walk = false; break; }
            node = first + c;
            level_base += 1 << (3 * l);
            at = at * 8 + int(digit);
            depth = l + 1;
        }
        (void)leaf_here;
        if (!walk) continue;
        D3 point = mk(0, 0, 0), normal = mk(0, 0, 0);
        __syncwarp();
        const int id = trace_front_warp(s, o, dir, point, normal, st_node[wib], st_mask[wib], node_tests, prim_tests, node,
                                        kMaxRefDepth - depth, 0, heavy.best + h, code);
        if (id >= 0 && lane == 0) { // publish: the lowest rank wins
            while (atomicCAS(heavy.lock + h, 0, 1) != 0) {}
            __threadfence();
            if (code < *reinterpret_cast<volatile unsigned*>(heavy.best + h)) {
                ids[lp] = id;
                if (points) {
                    points[lp] = point.x; points[map.n_local_pix + lp] = point.y; points[2 * (size_t)map.n_local_pix + lp] = point.z;
                    normals[lp] = normal.x; normals[map.n_local_pix + lp] = normal.y; normals[2 * (size_t)map.n_local_pix + lp] = normal.z;
                }
                __threadfence();
                *reinterpret_cast<volatile unsigned*>(heavy.best + h) = code;
            }
            __threadfence();
            atomicExch(heavy.lock + h, 0);
        }
        __syncwarp();
    }
    if (counters) {
        for (int off = 16; off > 0; off >>= 1) {
            node_tests += __shfl_xor_sync(0xffffffffu, node_tests, off);
            prim_tests += __shfl_xor_sync(0xffffffffu, prim_tests, off);
        }
        if (lane == 0) {
            atomicAdd(counters + 0, (unsigned long long)node_tests);
            atomicAdd(counters + 1, (unsigned long long)prim_tests);
            if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(counters + 5, (unsigned long long)n_heavy); // heavy rays of the frame
        }
    }
}

// ---- shading -----------------------------------------------------------------
constexpr double kRefPi = 3.1415926535; // entities.h:16

__device__ void texture_coord(const RefSceneD& s, const RefEntityD* __restrict__ e, D3 ip, int& ox, int& oy) {
    int x = 0, y = 0;
    switch (e->kind) {
    case 0:   // ImpSphere  entities.h:108-130
    case 4: { // ExpSphere  entities.h:549-571
        double radius = double(e->radius);
        double ulv = 2.0 * kRefPi * radius / 320.0;
        D3 ti = ip - ld3(e->pos);
        double cos_vert = dot3(ti, mk(0, 0, radius)) / (radius * radius);
        double atu = acos(cos_vert);
        if (e->kind == 0) y = ref_int((radius * atu) / ulv);
        else y = ref_int((0.5 * kRefPi * radius - radius * atu) / ulv);
        double small_r = radius * sin(atu);
        double cos_hori = dot3(mk(ti.x, ti.y, 0), mk(0, small_r, 0)) / (small_r * small_r);
        double ulh = 2.0 * kRefPi * small_r / 320.0;
        x = ref_int(small_r * acos(cos_hori) / ulh);
        break;
    }
    case 1: { // ImpTriangle entities.h:277-303
        const RefTriD* __restrict__ t = s.tris + e->tri_offset;
        D3 p1 = ld3(t->p1), p2 = ld3(t->p2), p3 = ld3(t->p3);
        D3 p2p1 = p2 - p1, p3p1 = p3 - p1, p3p2 = p3 - p2, ip1 = ip - p1;
        double p2p1_len = len3(p2p1), ip1_len = len3(ip1);
        double theta = acos(dot3(p2p1, ip1) / (p2p1_len * ip1_len));
        double ix_len = ip1_len * sin(theta);
        double v_len = len3(0.5 * (p2p1 + p3p1));
        double h_len = len3(0.5 * ((-p3p2) + (-p3p1)));
        double ulv = v_len / 160.0, ulh = h_len / 160.0;
        y = ref_int(ip1_len / ulh);
        x = ref_int(ix_len / ulv);
        break;
    }
    case 2: { // ExpRectangle entities.h:346-365 (aux0 = p3, aux1 = p4)
        const RefTriD* __restrict__ t = s.tris + e->tri_offset;
        D3 p1 = ld3(t->p1);
        D3 p3p1 = ld3(e->aux0) - p1, p4p1 = ld3(e->aux1) - p1;
        double width = len3(p4p1), length = len3(p3p1);
        double ulv = width / 64.0, ulh = length / 64.0;
        D3 ip1 = ip - p1;
        double ip1_len = len3(ip1);
        double cos_theta = acos(dot3(ip1, p3p1) / (length * ip1_len));
        x = ref_int(ip1_len * sin(acos(cos_theta)) / ulh);
        y = ref_int(ip1_len * cos_theta / ulv);
        break;
    }
    case 3: break; // ExpBox: (0,0)
    case 5: { // ExpQuad entities.h:630-641 (aux0 = vertices(0), aux1 = vertices(1))
        double width = double(e->f[0]), length = double(e->f[1]);
        double ulv = width / 160.0, ulh = length / 160.0;
        D3 right = ld3(e->aux0) - ld3(e->aux1);
        D3 ip1 = ip - ld3(e->aux1);
        double ip1_len = len3(ip1);
        double theta = acos(dot3(ip1, right) / (width * ip1_len));
        y = ref_int(ip1_len * sin(theta) / ulh);
        x = ref_int(ip1_len * cos(theta) / ulv);
        break;
    }
    case 6: { // ExpCube entities.h:769-811 (aux0 = vertices(0))
        double width = double(e->f[0]), length = double(e->f[1]);
        double ulv = width / 160.0, ulh = length / 160.0;
        D3 ipv = ip - ld3(e->aux0);
        double ip_len = len3(ipv);
        double theta = acos(dot3(ipv, mk(0, width, 0)) / (width * ip_len));
        y = ref_int(ip_len * sin(theta) / ulh);
        x = ref_int(ip_len * cos(theta) / ulv);
        break;
    }
    default: { // ExpCone entities.h:942-961 (f[0] = height, f[1] = radius)
        double radius = double(e->f[1]), height = double(e->f[0]);
        double ulh = sqrt(radius * radius + height * height) / 320.0;
        D3 pos = ld3(e->pos);
        double y_len = len3(ip - pos);
        y = ref_int(y_len / ulh);
        D3 centre = mk(double(float(pos.x)), double(float(pos.y)), double(float(ip.z)));
        double theta = atan(radius / height);
        double r_prime = y_len * sin(theta);
        D3 left = mk(0.0, double(float(r_prime)), 0.0);
        D3 ic = ip - centre;
        double ulv = 2.0 * kRefPi * r_prime / 320.0;
        double alpha = acos(dot3(ic, left) / (r_prime * r_prime));
        if (alpha > kRefPi / 4.0) alpha = acos(dot3(ic, -left) / (r_prime * r_prime));
        x = ref_int(r_prime * alpha / ulv);
        break;
    }
    }
    ox = x;
    oy = y;
}

// Material::blinn_phong_texture with the 32x32 checker evaluated in closed form
// instead of rebuilding the 12 KB pattern per pixel (material.h:49,66-92).
// pow(x, specular_power): the default exponent keeps its literal (the compiler may pick a different evaluation for a
// constant exponent, and the default path is the one pinned byte for byte against the reference's glibc pow)
__device__ __forceinline__ double spec_pow(double x, double power) { return power == 5.0 ? pow(x, 5.0) : pow(x, power); }

__device__ D3 blinn_phong_texture(const RefEntityD* __restrict__ e, D3 dir, D3 light, D3 ip, D3 normal, int u, int v) {
    const D3 color = ld3(e->color), shader = ld3(e->shader), spec = ld3(e->specular_color);
    int i = u % 32, j = v % 32;
    // negative remainders index the pattern block at a flat offset (see oracle/ref_restate.c)
    int flat = i * 32 + j;
    if (flat >= 0 && flat < 1024) { i = flat / 32; j = flat % 32; } else { i = 0; j = 0; }
    D3 tex;
    if ((i <= 16 && j <= 16) || (i > 16 && j > 16)) tex = mk(1, 1, 1);
    else tex = mk(double(ref_int(color.x)), double(ref_int(color.y)), double(ref_int(color.z)));
    D3 tdc = tex * 0.5;            // texture_diffuse_color (a literal 0.5 in material.h:51, not Material::diffuse_color)
    D3 la = tex * shader.x;
    D3 ldir = unit(light - ip);
    D3 ld = (std_max(0.0, dot3(normal, ldir)) * tdc) * shader.y;
    D3 bis = unit(unit(-dir) + unit(light - ip));
    D3 ls = (spec_pow(std_max(0.0, dot3(normal, bis)), e->specular_power) * spec) * shader.z;
    D3 out = la + ld + ls;
    return mk(std_min(out.x, 1.0), std_min(out.y, 1.0), std_min(out.z, 1.0));
}

// Material::blinn_phong, material.h:31-46 (unused by RayTracer::run, raytracer.h:79)
__device__ D3 blinn_phong_plain(const RefEntityD* __restrict__ e, D3 dir, D3 light, D3 ip, D3 normal) {
    const D3 color = ld3(e->color), shader = ld3(e->shader), spec = ld3(e->specular_color), diff = ld3(e->diffuse_color);
    D3 la = color * shader.x;
    D3 ldir = unit(light - ip);
    D3 ld = (std_max(0.0, dot3(normal, ldir)) * diff) * shader.y;
    D3 bis = unit(unit(-dir) + unit(light - ip));
    D3 ls = (spec_pow(std_max(0.0, dot3(normal, bis)), e->specular_power) * spec) * shader.z;
    D3 o = la + ld + ls;
    return mk(std_min(o.x, 1.0), std_min(o.y, 1.0), std_min(o.z, 1.0));
}

__global__ void __launch_bounds__(128) ref_shade_kernel(RefSceneD s, RefCamera cam, TileMap map, int lp0, int lp1,
                                                        const int32_t* __restrict__ ids,
                                                        const double* __restrict__ points,
                                                        const double* __restrict__ normals, uint8_t* __restrict__ rgb,
                                                        float* __restrict__ colour) {
    int lp = lp0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= lp1) return;
    int x, y;
    D3 c = mk(0, 0, 0);
    int id = ids[lp];
    if (id >= 0 && local_to_pixel(map, lp, x, y)) {
        size_t n = (size_t)map.n_local_pix;
        D3 ip = mk(points[lp], points[n + lp], points[2 * n + lp]);
        D3 nn = mk(normals[lp], normals[n + lp], normals[2 * n + lp]);
        D3 dir = pixel_dir(cam, x, y);
        const RefEntityD* __restrict__ e = s.entities + id;
        int u, v;
        texture_coord(s, e, ip, u, v);
        c = blinn_phong_texture(e, dir, ld3(cam.light), ip, nn, u, v);
    }
    // Image::setPixel: truncate, and QColor's validity rule (out of range -> black)
    int r = ref_int(255 * c.x), g = ref_int(255 * c.y), b = ref_int(255 * c.z);
    bool ok = r >= 0 && r <= 255 && g >= 0 && g <= 255 && b >= 0 && b <= 255;
    if (rgb) {
        rgb[3 * (size_t)lp + 0] = ok ? (uint8_t)r : 0;
        rgb[3 * (size_t)lp + 1] = ok ? (uint8_t)g : 0;
        rgb[3 * (size_t)lp + 2] = ok ? (uint8_t)b : 0;
    }
    if (colour) {
        colour[3 * (size_t)lp + 0] = float(c.x);
        colour[3 * (size_t)lp + 1] = float(c.y);
        colour[3 * (size_t)lp + 2] = float(c.z);
    }
}

// ---- probes --------------------------------------------------------------------
__global__ void probe_intersect_kernel(RefSceneD s, int entity, int n, const double* __restrict__ origins,
                                       const double* __restrict__ dirs, int32_t* __restrict__ hit,
                                       double* __restrict__ points, double* __restrict__ normals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 o = ld3(origins + 3 * i);
    D3 dir = unit(ld3(dirs + 3 * i)); // Ray ctor
    D3 p = mk(0, 0, 0), nn = mk(0, 0, 0);
    unsigned tests = 0;
    bool h = entity_hit(s, entity, o, dir, p, nn, tests);
    hit[i] = h ? 1 : 0;
    points[3 * i] = p.x; points[3 * i + 1] = p.y; points[3 * i + 2] = p.z;
    normals[3 * i] = nn.x; normals[3 * i + 1] = nn.y; normals[3 * i + 2] = nn.z;
}

// Forward DFS, full list (what Octree::intersect returns). One thread.
__global__ void probe_candidates_kernel(RefSceneD s, const double* __restrict__ od, int32_t* __restrict__ out,
                                        int max_out, int32_t* __restrict__ out_n) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    D3 o = ld3(od);
    D3 dir = unit(ld3(od + 3));
    int stack_node[kMaxRefDepth];
    signed char stack_next[kMaxRefDepth];
    int level = 0, n = 0;
    stack_node[0] = 0;
    stack_next[0] = 0;
    while (level >= 0) {
        const RefNodeD* nd = s.nodes + stack_node[level];
        int first = nd->first_child;
        if (first < 0) {
            for (int i = 0; i < nd->ent_count; ++i) {
                if (n < max_out) out[n] = s.ents[nd->ent_offset + i];
                ++n;
            }
            --level;
            continue;
        }
        int c = stack_next[level];
        bool pushed = false;
        for (; c < 8; ++c) {
            const RefNodeD* ch = s.nodes + first + c;
            if (ch->ent_count == 0) continue;
            if (node_box_hit(ch, o, dir)) {
                stack_next[level] = (signed char)(c + 1);
                if (level + 1 < kMaxRefDepth) {
                    ++level;
                    stack_node[level] = first + c;
                    stack_next[level] = 0;
                    pushed = true;
                }
                break;
            }
        }
        if (!pushed) --level;
    }
    *out_n = n;
}

__global__ void probe_texcoord_kernel(RefSceneD s, int entity, int n, const double* __restrict__ points,
                                      int32_t* __restrict__ uv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int u, v;
    texture_coord(s, s.entities + entity, ld3(points + 3 * i), u, v);
    uv[2 * i] = u;
    uv[2 * i + 1] = v;
}

// in15 = ray dir (normalised here, as the Ray ctor does), light, point, normal, unused
__global__ void probe_shade_kernel(RefSceneD s, int entity, int textured, const double* __restrict__ in, int u, int v,
                                   double* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    D3 dir = unit(ld3(in)), light = ld3(in + 3), ip = ld3(in + 6), nn = ld3(in + 9);
    const RefEntityD* __restrict__ e = s.entities + entity;
    D3 c = textured ? blinn_phong_texture(e, dir, light, ip, nn, u, v) : blinn_phong_plain(e, dir, light, ip, nn);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

// ---- untile ----------------------------------------------------------------------
__global__ void untile_kernel(TileMap map, const uint8_t* __restrict__ rgb_l, const int32_t* __restrict__ ids_l,
                              const float* __restrict__ rad_l, uint8_t* __restrict__ rgb_f,
                              int32_t* __restrict__ ids_f, float* __restrict__ rad_f) {
    int lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= map.n_local_pix) return;
    int x, y;
    if (!local_to_pixel(map, lp, x, y)) return;
    size_t g = (size_t)y * map.w + x;
    if (rgb_l && rgb_f) {
        rgb_f[3 * g] = rgb_l[3 * (size_t)lp];
        rgb_f[3 * g + 1] = rgb_l[3 * (size_t)lp + 1];
        rgb_f[3 * g + 2] = rgb_l[3 * (size_t)lp + 2];
    }
    if (ids_l && ids_f) ids_f[g] = ids_l[lp];
    if (rad_l && rad_f) {
        rad_f[3 * g] = rad_l[3 * (size_t)lp];
        rad_f[3 * g + 1] = rad_l[3 * (size_t)lp + 1];
        rad_f[3 * g + 2] = rad_l[3 * (size_t)lp + 2];
    }
}

inline int blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

} // namespace

void launch_ref_visibility(const RefSceneD& scene, const RefCamera& cam, const TileMap& map, int32_t* ids,
                           double* points, double* normals, unsigned long long* counters, cudaStream_t stream, int lp0,
                           int lp1, unsigned* next, const RefHeavyD* heavy) {
    if (lp1 < 0) lp1 = map.n_local_pix;
    if (lp1 <= lp0) return;
    if (scene.n_nodes > 1 && next) { // the tree has split: one warp per ray (node tests and leaf lists spread over the lanes)
        const int blocks = std::min(blocks_for(lp1 - lp0, kRefWarps), 148 * 16);
        cudaMemsetAsync(next, 0, sizeof(unsigned), stream);
        RefHeavyD hv{};
        if (heavy && heavy->budget > 0 && heavy->cap > 0) {
            hv = *heavy;
            cudaMemsetAsync(hv.count, 0, 2 * sizeof(unsigned), stream);
            cudaMemsetAsync(hv.best, 0xff, size_t(hv.cap) * sizeof(unsigned), stream);
            cudaMemsetAsync(hv.lock, 0, size_t(hv.cap) * sizeof(int), stream);
            cudaMemsetAsync(hv.masks, 0, size_t(hv.cap) * kHeavyCache * sizeof(unsigned short), stream);
        }
        ref_visibility_warp_kernel<<<blocks, kRefWarps * 32, 0, stream>>>(scene, cam, map, lp0, lp1, ids, points, normals, counters, next, hv);
        // the heavy rays of this band, each spread over many warps (a grid that finds the list empty leaves at once)
        if (hv.budget > 0) ref_heavy_kernel<<<148 * 4, kRefWarps * 32, 0, stream>>>(scene, cam, map, ids, points, normals, counters, hv);
        return;
    }
    ref_visibility_kernel<<<blocks_for(lp1 - lp0, 128), 128, 0, stream>>>(scene, cam, map, lp0, lp1, ids, points, normals,
                                                                           counters);
}

void launch_ref_shade(const RefSceneD& scene, const RefCamera& cam, const TileMap& map, const int32_t* ids,
                      const double* points, const double* normals, uint8_t* rgb, float* colour, cudaStream_t stream,
                      int lp0, int lp1) {
    if (lp1 < 0) lp1 = map.n_local_pix;
    if (lp1 <= lp0) return;
    ref_shade_kernel<<<blocks_for(lp1 - lp0, 128), 128, 0, stream>>>(scene, cam, map, lp0, lp1, ids, points, normals, rgb,
                                                                      colour);
}

void launch_probe_intersect(const RefSceneD& scene, int32_t entity, int n, const double* origins, const double* dirs,
                            int32_t* hit, double* points, double* normals, cudaStream_t stream) {
    if (n <= 0) return;
    probe_intersect_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(scene, entity, n, origins, dirs, hit, points,
                                                                   normals);
}

void launch_probe_candidates(const RefSceneD& scene, const double* origin_dir6, int32_t* out_ids, int max_out,
                             int32_t* out_n, cudaStream_t stream) {
    probe_candidates_kernel<<<1, 32, 0, stream>>>(scene, origin_dir6, out_ids, max_out, out_n);
}

void launch_probe_texcoord(const RefSceneD& scene, int32_t entity, int n, const double* points, int32_t* uv,
                           cudaStream_t stream) {
    if (n <= 0) return;
    probe_texcoord_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(scene, entity, n, points, uv);
}

void launch_probe_shade(const RefSceneD& scene, int32_t entity, int textured, const double* in15, int u, int v,
                        double* out_rgb, cudaStream_t stream) {
    probe_shade_kernel<<<1, 32, 0, stream>>>(scene, entity, textured, in15, u, v, out_rgb);
}

void launch_untile(const TileMap& map, const uint8_t* rgb_local, const int32_t* ids_local, const float* rad_local,
                   uint8_t* rgb_frame, int32_t* ids_frame, float* rad_frame, cudaStream_t stream) {
    if (map.n_local_pix == 0) return;
    untile_kernel<<<blocks_for(map.n_local_pix, 256), 256, 0, stream>>>(map, rgb_local, ids_local, rad_local, rgb_frame,
                                                                        ids_frame, rad_frame);
}

} // namespace g19
