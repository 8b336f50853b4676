// scene.cpp -- entity construction and the reference-semantics octree build.
//
// Reference regions mirrored here (host side of the path):
//   entity constructors   include/entities.h:45-47,138-148,310-324,381-406,
//                         461-506,581-590,652-727,823-899
//   boundingBox() quirks  include/entities.h:98-99,251-275,338-340,442,539-540,
//                         623-624,762-763,934-935 (in-class initialisers that run
//                         while Entity::pos is still {0,0,0})
//   Octree build          include/octree.h:20-30,75-129; include/bbox.h:33-35
// Compiled with -ffp-contract=off: the produced doubles must equal the
// reference's bit for bit (tests/test_scene_host.py checks them against the
// compiled reference).
#include "scene.h"

#include <cmath>
#include <cstring>
#include <new>

namespace g19 {
namespace {

constexpr double kRefPi = 3.1415926535; // entities.h:16

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) {
    double tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 unit(V3 v) { return v * (1.0 / std::sqrt(dot(v, v))); } // glm::normalize
inline V3 widen(float x, float y, float z) { return {double(x), double(y), double(z)}; }
inline V3 from(const double* p) { return {p[0], p[1], p[2]}; }
inline double lo(double a, double b) { return b < a ? b : a; }
inline double hi(double a, double b) { return a < b ? b : a; }

HostTri make_tri(V3 a, V3 b, V3 c) {
    HostTri t;
    t.p1 = a;
    t.p2 = b;
    t.p3 = c;
    t.edge1 = b - a;
    t.edge2 = c - a;
    t.normal = unit(cross(t.edge1, t.edge2));
    t.pos = 0.5 * (0.5 * (a + b) + c);
    return t;
}

// ImpTriangle::boundingBox (entities.h:251-275)
void tri_bounds(const HostTri& t, V3& mn, V3& mx) {
    mn = {lo(lo(t.p1.x, t.p2.x), t.p3.x), lo(lo(t.p1.y, t.p2.y), t.p3.y), lo(lo(t.p1.z, t.p2.z), t.p3.z)};
    V3 m = {hi(hi(t.p1.x, t.p2.x), t.p3.x), hi(hi(t.p1.y, t.p2.y), t.p3.y), hi(hi(t.p1.z, t.p2.z), t.p3.z) + 0.01};
    mx = m;
    if (m.x == mn.x) mx.x += 1e-5;
    if (m.y == mn.y) mx.y += 1e-5;
    if (m.z == mn.z) mx.z += 1e-5;
}

// BoundingBox(glm::vec3, glm::vec3): the corner values pass through float.
void float_bounds(HostEntity& e, double x0, double y0, double z0, double x1, double y1, double z1) {
    e.bbmin = widen(float(x0), float(y0), float(z0));
    e.bbmax = widen(float(x1), float(y1), float(z1));
}

// The two triangles of an ExpRectangle(p1,p2,p3). p4 is evaluated by an
// in-class initialiser while pos is still the origin, so p4 = 0 + (0 - p3).
void rect_pair(V3 p1, V3 p2, V3 p3, std::vector<HostTri>& out, V3* p4_out = nullptr) {
    V3 origin = {0, 0, 0};
    V3 p4 = origin + (origin - p3);
    out.push_back(make_tri(p1, p2, p3));
    out.push_back(make_tri(p1, p2, p4));
    if (p4_out) *p4_out = p4;
}

struct M3f { // glm::mat3, m[col][row]
    float m[3][3];
};
V3 apply(const M3f& r, V3 v) { // mat3 * vec3 with the dvec3 narrowed first
    float x = float(v.x), y = float(v.y), z = float(v.z);
    float ox = r.m[0][0] * x + r.m[1][0] * y + r.m[2][0] * z;
    float oy = r.m[0][1] * x + r.m[1][1] * y + r.m[2][1] * z;
    float oz = r.m[0][2] * x + r.m[1][2] * y + r.m[2][2] * z;
    return widen(ox, oy, oz);
}

} // namespace

void box_face_tris(V3 mn, V3 mx, std::vector<HostTri>& out) {
    V3 dlb = mn, drb = {mx.x, mn.y, mn.z}, dlt = {mn.x, mx.y, mn.z}, drt = {mx.x, mx.y, mn.z};
    V3 ulb = {mn.x, mn.y, mx.z}, urb = {mx.x, mn.y, mx.z}, ult = {mn.x, mx.y, mx.z}, urt = mx;
    rect_pair(dlb, urb, ulb, out);
    rect_pair(dlb, ult, dlt, out);
    rect_pair(dlb, drt, dlt, out);
    rect_pair(urt, ulb, ult, out);
    rect_pair(urt, drb, drt, out);
    rect_pair(urt, dlt, drt, out);
}

bool build_entity(const g19_entity_desc& d, HostEntity& e) {
    e = HostEntity();
    e.desc = d;
    e.kind = d.kind;
    e.first_tested = 0;
    e.radius = 0.f;
    e.in_tree = false;
    e.aux0 = e.aux1 = V3{0, 0, 0};
    const V3 pos = from(d.p);
    switch (d.kind) {
    case G19_IMP_SPHERE: {
        float r = d.f[0];
        e.combine = COMBINE_SPHERE;
        e.pos = pos;
        e.radius = r;
        float_bounds(e, 0.0 - r, 0.0 - r, 0.0 - r, 0.0 + r, 0.0 + r, 0.0 + r);
        return true;
    }
    case G19_IMP_TRIANGLE: {
        e.combine = COMBINE_SINGLE;
        e.tris.push_back(make_tri(from(d.p), from(d.p + 3), from(d.p + 6)));
        e.pos = e.tris[0].pos;
        tri_bounds(e.tris[0], e.bbmin, e.bbmax);
        return true;
    }
    case G19_EXP_RECTANGLE: {
        V3 p1 = from(d.p), p2 = from(d.p + 3), p3 = from(d.p + 6), p4;
        e.combine = COMBINE_FIRST;
        rect_pair(p1, p2, p3, e.tris, &p4);
        e.aux0 = p3;
        e.aux1 = p4;
        e.pos = 0.5 * (p1 + p2);
        e.bbmin = {lo(p1.x, p2.x), lo(p1.y, p2.y), lo(p1.z, p2.z)};
        e.bbmax = {hi(p1.x, p2.x), hi(p1.y, p2.y), hi(p1.z, p2.z)};
        return true;
    }
    case G19_EXP_BOX: {
        e.combine = COMBINE_BOX;
        box_face_tris(from(d.p), from(d.p + 3), e.tris);
        e.pos = {0, 0, 0};
        e.bbmin = from(d.p);
        e.bbmax = from(d.p + 3);
        return true;
    }
    case G19_EXP_SPHERE: {
        const int sectors = 10, stacks = 10;
        float radius = d.f[0];
        e.combine = COMBINE_NEAREST;
        e.first_tested = 1;
        e.pos = pos;
        e.radius = radius;
        float sector_step = float(2 * kRefPi / sectors);
        float stack_step = float(kRefPi / stacks);
        std::vector<V3> grid;
        grid.reserve((stacks + 1) * (sectors + 1));
        for (int i = 0; i <= stacks; ++i) {
            float stack_angle = float(kRefPi / 2 - i * stack_step);
            float ring = radius * cosf(stack_angle);
            float z = float(radius * sinf(stack_angle) - pos.z);
            for (int j = 0; j <= sectors; ++j) {
                float sector_angle = j * sector_step;
                float x = float(ring * cosf(sector_angle) - pos.x);
                float y = float(ring * sinf(sector_angle) - pos.y);
                grid.push_back(widen(x, y, z));
            }
        }
        for (int i = 0; i < stacks; ++i) {
            int k1 = i * (sectors + 1), k2 = k1 + sectors + 1;
            for (int j = 0; j < sectors; ++j, ++k1, ++k2) {
                if (i != 0) e.tris.push_back(make_tri(grid[k1], grid[k2], grid[k1 + 1]));
                if (i != stacks - 1) e.tris.push_back(make_tri(grid[k1 + 1], grid[k2], grid[k2 + 1]));
            }
        }
        float_bounds(e, 0.0 - radius, 0.0 - radius, 0.0 - radius, 0.0 + radius, 0.0 + radius, 0.0 + radius);
        return true;
    }
    case G19_EXP_QUAD: {
        float width = d.f[0], length = d.f[1], alpha = d.f[2];
        e.combine = COMBINE_NEAREST;
        e.pos = pos;
        // cos(float)/sin(float) pick the float overloads in the reference
        float ca = cosf(alpha), sa = sinf(alpha);
        double xr = pos.x + width / 2, xl = pos.x - width / 2;
        V3 ur = {xr * ca, pos.y + length / 2, pos.z + xr * sa};
        V3 ul = {xl * ca, pos.y + length / 2, pos.z + xl * sa};
        V3 dr = {xr * ca, pos.y - length / 2, pos.z + xr * sa};
        V3 dl = {xl * ca, pos.y - length / 2, pos.z + pos.z + xl * sa}; // the doubled pos.z is the reference's
        e.tris.push_back(make_tri(ul, dr, ur));
        e.tris.push_back(make_tri(ul, dl, dr));
        e.aux0 = ur;
        e.aux1 = ul;
        float_bounds(e, 0.0 - width / 2, 0.0 - length / 2, 0.0, 0.0 + width / 2, 0.0 + length / 2,
                     0.0 + (0.0 + width / 2) * sa);
        return true;
    }
    case G19_EXP_CUBE: {
        float w = d.f[0], l = d.f[1], h = d.f[2];
        e.combine = COMBINE_NEAREST;
        e.pos = pos;
        double x0 = pos.x - w / 2, x1 = pos.x + w / 2, y0 = pos.y - l / 2, y1 = pos.y + l / 2, z0 = pos.z - h / 2,
               z1 = pos.z + h / 2;
        const V3 c[8] = {{x0, y0, z0}, {x0, y0, z1}, {x1, y0, z0}, {x1, y0, z1},
                         {x0, y1, z1}, {x0, y1, z0}, {x1, y1, z0}, {x1, y1, z1}};
        static const int8_t faces[12][3] = {{0, 1, 2}, {3, 1, 2}, {4, 5, 7}, {7, 5, 6}, {1, 0, 4}, {4, 0, 5},
                                            {3, 7, 2}, {7, 6, 2}, {1, 4, 3}, {3, 4, 7}, {0, 5, 2}, {2, 5, 6}};
        for (auto& f : faces) e.tris.push_back(make_tri(c[f[0]], c[f[1]], c[f[2]]));
        e.aux0 = c[0];
        float_bounds(e, 0.0 - w / 2, 0.0 - l / 2, 0.0 - h / 2, 0.0 + w / 2, 0.0 + l / 2, 0.0 + h / 2);
        return true;
    }
    case G19_EXP_CONE: {
        float height = d.f[0], radius = d.f[1];
        e.combine = COMBINE_NEAREST;
        e.pos = pos;
        e.radius = radius;
        // The reference overwrites its own ctor parameter (entities.h:825): the
        // axis is always normalize({-1,0,-10}), whatever the caller passed.
        V3 axis = unit(V3{-1, 0, -10});
        M3f rot_x = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}, rot_y = rot_x;
        V3 yz = {0, axis.y, axis.z};
        if (!(yz.x == 0 && yz.y == 0 && yz.z == 0)) {
            double a = ((axis.y < 0) ? 1.0 : -1.0) * std::acos(dot(unit(yz), V3{0, 0, -1}));
            rot_x = {{{1, 0, 0}, {0, float(std::cos(a)), float(-std::sin(a))}, {0, float(std::sin(a)), float(std::cos(a))}}};
        }
        V3 xz = {axis.x, 0, -std::sqrt(axis.z * axis.z + axis.y * axis.y)};
        if (!(xz.x == 0 && xz.y == 0 && xz.z == 0)) {
            double a = ((axis.x > 0) ? 1.0 : -1.0) * std::acos(dot(unit(xz), V3{0, 0, -1}));
            rot_y = {{{float(std::cos(a)), 0, float(std::sin(a))}, {0, 1, 0}, {float(-std::sin(a)), 0, float(std::cos(a))}}};
        }
        std::vector<V3> rim;
        const double subdivisions = 50.0;
        for (int i = 0; i <= subdivisions; ++i) {
            float deg = float(i * 360.0 / subdivisions);
            V3 p = {pos.x + radius * std::cos(deg * kRefPi / 180.0), pos.y + radius * std::sin(deg * kRefPi / 180.0),
                    pos.z - height};
            p = p - pos;
            p = apply(rot_x, p);
            p = apply(rot_y, p);
            rim.push_back(p + pos);
        }
        V3 base_centre = pos + unit(axis) * double(height);
        for (size_t i = 0; i + 1 < rim.size(); ++i) {
            e.tris.push_back(make_tri(pos, rim[i], rim[i + 1]));
            e.tris.push_back(make_tri(base_centre, rim[i], rim[i + 1]));
        }
        float_bounds(e, 0.0 - radius, 0.0 - radius, 0.0 - height, 0.0 + radius, 0.0 + radius, 0.0);
        return true;
    }
    default: return false;
    }
}

// ---- Octree build --------------------------------------------------------
namespace {

bool overlaps(V3 amin, V3 amax, V3 bmin, V3 bmax) { // BoundingBox::intersect, bbox.h:25-39
    V3 ca = 0.5 * (amin + amax), cb = 0.5 * (bmin + bmax);
    V3 gap = ca - cb;
    bool x = std::fabs(gap.x) < (0.5 * (amax.x - amin.x) + 0.5 * (bmax.x - bmin.x));
    bool y = std::fabs(gap.y) < (0.5 * (amax.y - amin.y) + 0.5 * (bmax.y - bmin.y));
    bool z = std::fabs(gap.z) < (0.5 * (amax.z - amin.z) + 0.5 * (bmax.z - bmin.z));
    return x && y && z;
}
bool le3(V3 a, V3 b) { return a.x <= b.x && a.y <= b.y && a.z <= b.z; }

// Node::partition (octree.h:75-110): split a leaf unless every stored entity
// straddles the centre. Entities already stored are NOT redistributed.
void maybe_split(g19_scene& s, int32_t ni) {
    if (s.nodes[ni].first_child >= 0) return;
    V3 a = s.nodes[ni].mn, b = s.nodes[ni].mx;
    V3 m = (a + b) * 0.5;
    bool straddle_all = true;
    for (int32_t ei : s.nodes[ni].ents) {
        const HostEntity& e = s.ents[ei];
        straddle_all = straddle_all && le3(e.bbmin, m) && le3(m, e.bbmax);
    }
    if (straddle_all) return;
    int32_t base = int32_t(s.nodes.size());
    const V3 lows[8] = {a, {a.x, m.y, a.z}, {m.x, a.y, a.z}, {m.x, m.y, a.z}, m, {a.x, m.y, m.z}, {m.x, a.y, m.z}, {a.x, a.y, m.z}};
    const V3 highs[8] = {m, {m.x, b.y, m.z}, {b.x, m.y, m.z}, {b.x, b.y, m.z}, b, {m.x, b.y, b.z}, {b.x, m.y, b.z}, {m.x, m.y, b.z}};
    for (int c = 0; c < 8; ++c) {
        HostNode child;
        child.mn = lows[c];
        child.mx = highs[c];
        child.first_child = -1;
        s.nodes.push_back(std::move(child));
    }
    s.nodes[ni].first_child = base;
}

void descend(g19_scene& s, int32_t ni, int32_t ei) { // Node::push_obj, octree.h:115-129
    s.nodes[ni].ents.push_back(ei);
    maybe_split(s, ni);
    int32_t base = s.nodes[ni].first_child;
    if (base < 0) return;
    const V3 emin = s.ents[ei].bbmin, emax = s.ents[ei].bbmax;
    for (int c = 0; c < 8; ++c) {
        V3 cmin = s.nodes[base + c].mn, cmax = s.nodes[base + c].mx;
        if (le3(cmin, emin) && le3(emax, cmax)) descend(s, base + c, ei);
        else if (overlaps(cmin, cmax, emin, emax)) s.nodes[base + c].ents.push_back(ei);
    }
}

int depth_of(const g19_scene& s, int32_t ni) {
    int32_t base = s.nodes[ni].first_child;
    if (base < 0) return 0;
    int d = 0;
    for (int c = 0; c < 8; ++c) {
        int k = depth_of(s, base + c);
        if (k > d) d = k;
    }
    return d + 1;
}

} // namespace

bool push_back(g19_scene& s, int32_t ei) {
    HostEntity& e = s.ents[ei];
    if (!overlaps(s.nodes[0].mn, s.nodes[0].mx, e.bbmin, e.bbmax)) return false; // octree.h:22-24, silent
    e.in_tree = true;
    descend(s, 0, ei);
    return true;
}

int max_depth(const g19_scene& s) { return depth_of(s, 0); }

} // namespace g19

// ---- C ABI: scene ------------------------------------------------------------
using namespace g19;

extern "C" {

int g19_scene_create(const double mn[3], const double mx[3], g19_scene** out) {
    if (!mn || !mx || !out) return G19_ERR_INVALID;
    g19_scene* s = new (std::nothrow) g19_scene();
    if (!s) return G19_ERR_INVALID;
    s->rmin = {mn[0], mn[1], mn[2]};
    s->rmax = {mx[0], mx[1], mx[2]};
    HostNode root;
    root.mn = s->rmin;
    root.mx = s->rmax;
    root.first_child = -1;
    s->nodes.push_back(std::move(root));
    *out = s;
    return G19_OK;
}

void g19_scene_destroy(g19_scene* s) { delete s; }

int g19_scene_add_entity(g19_scene* s, const g19_entity_desc* d, int32_t* out_index) {
    if (!s || !d) return G19_ERR_INVALID;
    HostEntity e;
    if (!build_entity(*d, e)) return G19_ERR_INVALID;
    int32_t idx = int32_t(s->ents.size());
    s->ents.push_back(std::move(e));
    if (out_index) *out_index = idx;
    return push_back(*s, idx) ? G19_OK : G19_ERR_REJECTED;
}

int g19_scene_entity_count(const g19_scene* s) { return s ? int(s->ents.size()) : 0; }

int g19_scene_set_shapes(g19_scene* s, int shapes) {
    if (!s || (shapes != G19_SHAPES_REF && shapes != G19_SHAPES_FIXED)) return G19_ERR_INVALID;
    s->shapes = shapes;
    return G19_OK;
}

int g19_scene_get_entity(const g19_scene* s, int32_t i, g19_entity_desc* out) {
    if (!s || !out || i < 0 || size_t(i) >= s->ents.size()) return G19_ERR_INVALID;
    *out = s->ents[i].desc;
    return G19_OK;
}

int g19_scene_entity_bbox(const g19_scene* s, int32_t i, double o[6]) {
    if (!s || !o || i < 0 || size_t(i) >= s->ents.size()) return G19_ERR_INVALID;
    const HostEntity& e = s->ents[i];
    o[0] = e.bbmin.x; o[1] = e.bbmin.y; o[2] = e.bbmin.z;
    o[3] = e.bbmax.x; o[4] = e.bbmax.y; o[5] = e.bbmax.z;
    return G19_OK;
}

int g19_scene_entity_triangles(const g19_scene* s, int32_t i, double* out, int max_tris) {
    if (!s || i < 0 || size_t(i) >= s->ents.size()) return 0;
    const HostEntity& e = s->ents[i];
    int n = int(e.tris.size());
    for (int k = 0; k < n && k < max_tris && out; ++k) {
        const HostTri& t = e.tris[k];
        double v[9] = {t.p1.x, t.p1.y, t.p1.z, t.p2.x, t.p2.y, t.p2.z, t.p3.x, t.p3.y, t.p3.z};
        std::memcpy(out + 9 * k, v, sizeof v);
    }
    return n;
}

int g19_entity_bbox(const g19_entity_desc* d, double o[6]) {
    HostEntity e;
    if (!d || !o || !build_entity(*d, e)) return G19_ERR_INVALID;
    o[0] = e.bbmin.x; o[1] = e.bbmin.y; o[2] = e.bbmin.z;
    o[3] = e.bbmax.x; o[4] = e.bbmax.y; o[5] = e.bbmax.z;
    return G19_OK;
}

int g19_entity_triangles(const g19_entity_desc* d, double* out, int max_tris) {
    HostEntity e;
    if (!d || !build_entity(*d, e)) return 0;
    int n = int(e.tris.size());
    for (int k = 0; k < n && k < max_tris && out; ++k) {
        const HostTri& t = e.tris[k];
        double v[9] = {t.p1.x, t.p1.y, t.p1.z, t.p2.x, t.p2.y, t.p2.z, t.p3.x, t.p3.y, t.p3.z};
        std::memcpy(out + 9 * k, v, sizeof v);
    }
    return n;
}

int g19_scene_builtin(int which, int n, int w, int h, g19_scene** out, g19_camera* cam, double light[3]) {
    return make_builtin(which, n, w, h, out, cam, light);
}

} // extern "C"
