// fixed_shapes.cpp -- G19_SHAPES_FIXED: the composite entities as their constructors MEANT to build them.
//
// SURVEY.md 8(f) row 4 / hard part 1: several reference constructors contain bugs that define what REF mode shows
// (and REF mode keeps them, bit for bit, scene.cpp). PATH mode bounces light off geometry, and light bouncing off a
// sphere tessellated around -pos or a rectangle whose fourth corner is -p3 is not a meaningful picture. With
// g19_scene_set_shapes(scene, G19_SHAPES_FIXED) PATH mode's primitive extraction uses the triangles below instead --
// same entity ids, same materials; REF mode is untouched.
//
//   ExpRectangle(p1,p2,p3)  reference include/entities.h:308-324: p4 is initialised from this->pos before the
//                           constructor body sets pos (:319 vs :312), so p4 = -p3. Fixed: p4 = p1 + p2 - p3.
//   ExpBox(min,max)         entities.h:379-406: six ExpRectangle faces, each with that p4. Fixed: proper faces.
//   ExpSphere(pos,r)        entities.h:457-506: vertices are offset by -pos (:475,481-482) and intersect() skips
//                           triangle 0 (:520). Fixed: pos + r (cos phi cos theta, cos phi sin theta, sin phi), the
//                           same 10 x 10 stack/sector triangulation, every triangle tested.
//   ExpQuad(pos,w,l,alpha)  entities.h:579-590: rotates (pos.x +- w/2) about the ORIGIN and adds pos.z twice to the
//                           fourth vertex (:586). Fixed: a w x l quad centred on pos, width along (cos a, 0, sin a).
//   ExpCone(pos,dir,h,r)    entities.h:821-899: overwrites the caller's dir with normalize(-1,0,-10) (:825).
//                           Fixed: apex pos, base circle of 50 segments at pos + h * normalize(dir), the caller's dir.
//   ExpCube, ImpTriangle, ImpSphere are geometrically right in the reference: unchanged (false is returned).
// The same formulas, in the same operation order, live in oracle/path_oracle.c (fixed_tris) so that both sides
// round to the same float vertices.
#include <cmath>

#include "scene.h"

namespace g19 {
namespace {

HostTri tri3(V3 a, V3 b, V3 c) {
    HostTri t{};
    t.p1 = a;
    t.p2 = b;
    t.p3 = c;
    return t;
}
V3 add3(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 sub3(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 mul3(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
V3 norm3(V3 a) {
    double l = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return l > 0 ? mul3(a, 1.0 / l) : a;
}
void rect(V3 p1, V3 p2, V3 p3, std::vector<HostTri>& out) {
    V3 p4 = sub3(add3(p1, p2), p3);
    out.push_back(tri3(p1, p2, p3));
    out.push_back(tri3(p1, p2, p4));
}

} // namespace

bool fixed_triangles(const g19_entity_desc& d, std::vector<HostTri>& out) {
    const double kPi = 3.14159265358979323846;
    const V3 pos = {d.p[0], d.p[1], d.p[2]};
    switch (d.kind) {
    case G19_EXP_RECTANGLE:
        rect(pos, V3{d.p[3], d.p[4], d.p[5]}, V3{d.p[6], d.p[7], d.p[8]}, out);
        return true;
    case G19_EXP_BOX: {
        const V3 mn = pos, mx = {d.p[3], d.p[4], d.p[5]};
        // one rectangle per face: diagonal corners first, a third corner last
        rect(V3{mn.x, mn.y, mn.z}, V3{mx.x, mx.y, mn.z}, V3{mx.x, mn.y, mn.z}, out); // z = min
        rect(V3{mn.x, mn.y, mx.z}, V3{mx.x, mx.y, mx.z}, V3{mx.x, mn.y, mx.z}, out); // z = max
        rect(V3{mn.x, mn.y, mn.z}, V3{mx.x, mn.y, mx.z}, V3{mx.x, mn.y, mn.z}, out); // y = min
        rect(V3{mn.x, mx.y, mn.z}, V3{mx.x, mx.y, mx.z}, V3{mx.x, mx.y, mn.z}, out); // y = max
        rect(V3{mn.x, mn.y, mn.z}, V3{mn.x, mx.y, mx.z}, V3{mn.x, mx.y, mn.z}, out); // x = min
        rect(V3{mx.x, mn.y, mn.z}, V3{mx.x, mx.y, mx.z}, V3{mx.x, mx.y, mn.z}, out); // x = max
        return true;
    }
    case G19_EXP_SPHERE: {
        const double r = double(d.f[0]);
        const int stacks = 10, sectors = 10;
        std::vector<V3> v;
        for (int i = 0; i <= stacks; ++i) {
            const double phi = kPi / 2 - double(i) * (kPi / stacks);
            const double ring = r * std::cos(phi), z = r * std::sin(phi);
            for (int j = 0; j <= sectors; ++j) {
                const double theta = double(j) * (2 * kPi / sectors);
                v.push_back(V3{pos.x + ring * std::cos(theta), pos.y + ring * std::sin(theta), pos.z + z});
            }
        }
        for (int i = 0; i < stacks; ++i) {
            int k1 = i * (sectors + 1), k2 = k1 + sectors + 1;
            for (int j = 0; j < sectors; ++j, ++k1, ++k2) {
                if (i != 0) out.push_back(tri3(v[k1], v[k2], v[k1 + 1]));
                if (i != stacks - 1) out.push_back(tri3(v[k1 + 1], v[k2], v[k2 + 1]));
            }
        }
        return true;
    }
    case G19_EXP_QUAD: {
        const double hw = double(d.f[0]) / 2, hl = double(d.f[1]) / 2, a = double(d.f[2]);
        const double c = std::cos(a), s = std::sin(a);
        const V3 v0 = {pos.x + hw * c, pos.y + hl, pos.z + hw * s}; // up right
        const V3 v1 = {pos.x - hw * c, pos.y + hl, pos.z - hw * s}; // up left
        const V3 v2 = {pos.x + hw * c, pos.y - hl, pos.z + hw * s}; // down right
        const V3 v3 = {pos.x - hw * c, pos.y - hl, pos.z - hw * s}; // down left
        out.push_back(tri3(v1, v2, v0));
        out.push_back(tri3(v1, v3, v2));
        return true;
    }
    case G19_EXP_CONE: {
        const double h = double(d.f[0]), r = double(d.f[1]);
        V3 axis = norm3(V3{d.p[3], d.p[4], d.p[5]});
        if (axis.x == 0 && axis.y == 0 && axis.z == 0) axis = V3{0, 0, -1};
        const V3 centre = add3(pos, mul3(axis, h));
        const V3 helper = std::fabs(axis.x) < 0.9 ? V3{1, 0, 0} : V3{0, 1, 0};
        const V3 u = norm3(cross3(axis, helper)), w = cross3(axis, u);
        const int n = 50;
        std::vector<V3> rim;
        for (int i = 0; i <= n; ++i) {
            const double ang = double(i) * (2 * kPi / n);
            const double cu = r * std::cos(ang), cw = r * std::sin(ang);
            rim.push_back(V3{centre.x + cu * u.x + cw * w.x, centre.y + cu * u.y + cw * w.y, centre.z + cu * u.z + cw * w.z});
        }
        for (int i = 0; i < n; ++i) {
            out.push_back(tri3(pos, rim[i], rim[i + 1]));
            out.push_back(tri3(centre, rim[i], rim[i + 1]));
        }
        return true;
    }
    default: return false;
    }
}

} // namespace g19
