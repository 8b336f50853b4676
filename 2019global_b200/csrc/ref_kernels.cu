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

// ImpTriangle(p1,p2,p3) derived members, entities.h:138-148
__device__ __forceinline__ Tri tri_make(D3 p1, D3 p2, D3 p3) {
    Tri r;
    r.p1 = p1; r.p2 = p2; r.p3 = p3;
    D3 e1 = p2 - p1, e2 = p3 - p1;
    r.n = unit(cross3(e1, e2));
    r.pos = 0.5 * (0.5 * (p1 + p2) + p3);
    r.e1x = float(e1.x); r.e1y = float(e1.y); r.e1z = float(e1.z);
    r.e2x = float(e2.x); r.e2y = float(e2.y); r.e2z = float(e2.z);
    return r;
}

// glm::length(d) < 1e-3 without paying for the sqrt in the common case.
__device__ __forceinline__ bool shorter_than_eps(D3 d) {
    double q = dot3(d, d);
    return (q < 1.0e-5) && (sqrt(q) < 1.0e-3);
}

// ImpTriangle::intersect, entities.h:150-249. `facing` = normal as returned.
__device__ __forceinline__ bool tri_hit(const Tri& t, D3 o, D3 dir, D3& point, D3& facing) {
    double nd = dot3(t.n, dir);
    if (nd == 0) return false;
    // A = transpose(mat3(edge1, edge2, -dir)) in float; only row 2 of inverse(A) is consumed
    float m02 = float(-dir.x), m12 = float(-dir.y), m22 = float(-dir.z);
    float m00 = t.e1x, m10 = t.e1y, m20 = t.e1z;
    float m01 = t.e2x, m11 = t.e2y, m21 = t.e2z;
    float det = +m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02);
    float ood = 1.0f / det;
    float i20 = +(m10 * m21 - m20 * m11) * ood;
    float i21 = -(m00 * m21 - m20 * m01) * ood;
    float i22 = +(m00 * m11 - m10 * m01) * ood;
    D3 rhs = o - t.pos;
    float vx = float(rhs.x), vy = float(rhs.y), vz = float(rhs.z);
    float solz = i20 * vx + i21 * vy + i22 * vz;
    D3 p = o + double(solz) * dir;

    D3 d1 = unit(cross3(t.p1 - p, t.p2 - p));
    D3 d2 = unit(cross3(t.p2 - p, t.p3 - p));
    D3 d3 = unit(cross3(t.p3 - p, t.p1 - p));
    bool hit;
    if (shorter_than_eps(d1) || shorter_than_eps(d2) || shorter_than_eps(d3)) {
        hit = true;
    } else {
        D3 a = d1 - d2, b = d2 - d3;
        bool c1 = sqr(a.x) + sqr(a.y) + sqr(a.z) < 1.0e-3;
        bool c2 = sqr(b.x) + sqr(b.y) + sqr(b.z) < 1.0e-3;
        hit = c1 && c2;
    }
    if (!hit) return false;
    point = p;
    facing = (nd < 0) ? t.n : -t.n; // dot(ray.dir, normal) < 0 -- same products, same sum
    return true;
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
__device__ bool node_box_hit(const RefNodeD* __restrict__ nd, D3 o, D3 dir) {
    const D3 mn = ld3(nd->mn), mx = ld3(nd->mx);
    const D3 origin = mk(0, 0, 0);
#pragma unroll 1
    for (int f = 0; f < 6; ++f) {
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
        D3 pt, nn;
        Tri t = tri_make(p1, p2, p3);
        if (tri_hit(t, o, dir, pt, nn)) return true;
        D3 p4 = origin + (origin - p3);
        t = tri_make(p1, p2, p4);
        if (tri_hit(t, o, dir, pt, nn)) return true;
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

__global__ void __launch_bounds__(128) ref_visibility_kernel(RefSceneD s, RefCamera cam, TileMap map,
                                                             int32_t* __restrict__ ids, double* __restrict__ points,
                                                             double* __restrict__ normals,
                                                             unsigned long long* __restrict__ counters) {
    int lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= map.n_local_pix) return;
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
__device__ D3 blinn_phong_texture(D3 color, D3 dir, D3 light, D3 ip, D3 normal, int u, int v) {
    int i = u % 32, j = v % 32;
    // negative remainders index the pattern block at a flat offset (see oracle/ref_restate.c)
    int flat = i * 32 + j;
    if (flat >= 0 && flat < 1024) { i = flat / 32; j = flat % 32; } else { i = 0; j = 0; }
    D3 tex;
    if ((i <= 16 && j <= 16) || (i > 16 && j > 16)) tex = mk(1, 1, 1);
    else tex = mk(double(ref_int(color.x)), double(ref_int(color.y)), double(ref_int(color.z)));
    D3 tdc = tex * 0.5;
    D3 la = tex * 0.1;
    D3 ldir = unit(light - ip);
    D3 ld = (std_max(0.0, dot3(normal, ldir)) * tdc) * 0.7;
    D3 bis = unit(unit(-dir) + unit(light - ip));
    D3 ls = (pow(std_max(0.0, dot3(normal, bis)), 5.0) * mk(1, 1, 1)) * 1.0;
    D3 out = la + ld + ls;
    return mk(std_min(out.x, 1.0), std_min(out.y, 1.0), std_min(out.z, 1.0));
}

__global__ void __launch_bounds__(128) ref_shade_kernel(RefSceneD s, RefCamera cam, TileMap map,
                                                        const int32_t* __restrict__ ids,
                                                        const double* __restrict__ points,
                                                        const double* __restrict__ normals, uint8_t* __restrict__ rgb,
                                                        float* __restrict__ colour) {
    int lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= map.n_local_pix) return;
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
        c = blinn_phong_texture(ld3(e->color), dir, ld3(cam.light), ip, nn, u, v);
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
    D3 color = ld3(s.entities[entity].color);
    D3 c;
    if (textured) {
        c = blinn_phong_texture(color, dir, light, ip, nn, u, v);
    } else { // Material::blinn_phong, material.h:31-46 (unused by RayTracer::run, raytracer.h:79)
        D3 la = color * 0.1;
        D3 ldir = unit(light - ip);
        D3 ld = (std_max(0.0, dot3(nn, ldir)) * (color * 0.5)) * 0.7;
        D3 bis = unit(unit(-dir) + unit(light - ip));
        D3 ls = (pow(std_max(0.0, dot3(nn, bis)), 5.0) * mk(1, 1, 1)) * 1.0;
        D3 o = la + ld + ls;
        c = mk(std_min(o.x, 1.0), std_min(o.y, 1.0), std_min(o.z, 1.0));
    }
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
                           double* points, double* normals, unsigned long long* counters, cudaStream_t stream) {
    if (map.n_local_pix == 0) return;
    ref_visibility_kernel<<<blocks_for(map.n_local_pix, 128), 128, 0, stream>>>(scene, cam, map, ids, points, normals,
                                                                                 counters);
}

void launch_ref_shade(const RefSceneD& scene, const RefCamera& cam, const TileMap& map, const int32_t* ids,
                      const double* points, const double* normals, uint8_t* rgb, float* colour, cudaStream_t stream) {
    if (map.n_local_pix == 0) return;
    ref_shade_kernel<<<blocks_for(map.n_local_pix, 128), 128, 0, stream>>>(scene, cam, map, ids, points, normals, rgb,
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
