// builtin_scenes.cpp -- the procedural scenes of BASELINE.json `configs`.
//
// One source of truth: each generator emits plain g19_entity_desc records, in
// push order, through the same g19_scene_add_entity path a caller would use.
// The tests read them back (g19_scene_get_entity) to build the identical scene
// on the compiled reference and on the oracle.
//
//   G19_SCENE_DEFAULT        the literal of reference main.cpp:24-57 (config 1)
//   G19_SCENE_CORNELL        SURVEY.md 8(d) C2/C5: room x in [-12,8], y,z in
//                            [-6,6]; 5 walls x 2 ImpTriangle pushed FIRST, the
//                            ceiling light next, 2 ImpSphere r=2 pushed LAST
//                            (far-to-near, so the reference's last-hit-wins
//                            selection coincides with nearest-hit)
//   G19_SCENE_CORNELL_GLASS  C3: same, one mirror and one glass (ior 1.5) sphere
//   G19_SCENE_HEIGHTFIELD    C4: x = 6 + 1.5 sin(0.7y) cos(0.9z) over y,z in
//                            [-8,8], n x n cells, 2 n^2 ImpTriangle entities
//                            (n = 708 -> 1 002 528 triangles), lit by a panel behind the camera
//   G19_SCENE_HEIGHTFIELD_ROOM  C4 as benchmarked: the same surface as the far wall of a closed
//                            room with a ceiling light, so that paths live to the depth limit
#include <cmath>

#include "scene.h"

namespace g19 {
namespace {

g19_entity_desc blank(int kind) {
    g19_entity_desc d = {};
    d.kind = kind;
    d.bsdf = G19_BSDF_DIFFUSE;
    d.color[0] = 1.0; // Entity() default material, entities.h:21
    d.ior = 1.5f;
    return d;
}

void set3(double* dst, double x, double y, double z) {
    dst[0] = x;
    dst[1] = y;
    dst[2] = z;
}

void add(g19_scene* s, const g19_entity_desc& d) { g19_scene_add_entity(s, &d, nullptr); }

void add_triangle(g19_scene* s, const double a[3], const double b[3], const double c[3], const double color[3],
                  int bsdf = G19_BSDF_DIFFUSE, float emit = 0.f) {
    g19_entity_desc d = blank(G19_IMP_TRIANGLE);
    set3(d.p, a[0], a[1], a[2]);
    set3(d.p + 3, b[0], b[1], b[2]);
    set3(d.p + 6, c[0], c[1], c[2]);
    set3(d.color, color[0], color[1], color[2]);
    d.bsdf = bsdf;
    d.emission[0] = d.emission[1] = d.emission[2] = emit;
    add(s, d);
}

// a planar quad a-b-c-d as two triangles (a,b,c) and (a,c,d)
void add_quad(g19_scene* s, const double a[3], const double b[3], const double c[3], const double d[3],
              const double color[3], int bsdf = G19_BSDF_DIFFUSE, float emit = 0.f) {
    add_triangle(s, a, b, c, color, bsdf, emit);
    add_triangle(s, a, c, d, color, bsdf, emit);
}

void add_sphere(g19_scene* s, double x, double y, double z, float r, const double color[3], int bsdf) {
    g19_entity_desc d = blank(G19_IMP_SPHERE);
    set3(d.p, x, y, z);
    d.f[0] = r;
    set3(d.color, color[0], color[1], color[2]);
    d.bsdf = bsdf;
    add(s, d);
}

// The reference camera uses the image WIDTH for the vertical extent as well
// (raytracer.h:30), so in a non-square frame the optical axis is off-centre.
// `up` is never orthogonalised (camera.h:8), so pitching lookAt down by theta
// with focal*sin(theta) = (w-h)/2 * 0.0002 puts the image plane back in the
// vertical plane, centred -- using only the public Camera(pos, lookAt, focal).
void centred_camera(g19_camera* cam, double px, double py, double pz, double focal, int w, int h) {
    if (!cam) return;
    double s = (double(w) - double(h)) * 0.5 * 0.0002 / focal;
    if (s > 0.95) s = 0.95;
    if (s < -0.95) s = -0.95;
    double c = std::sqrt(1.0 - s * s);
    set3(cam->pos, px, py, pz);
    set3(cam->look_at, px + c, py, pz - s);
    cam->focal = focal;
}

int new_scene(g19_scene** out, double half) {
    double mn[3] = {-half, -half, -half}, mx[3] = {half, half, half};
    return g19_scene_create(mn, mx, out);
}

void cornell(g19_scene* s, bool glass) {
    const double white[3] = {0.73, 0.73, 0.73}, red[3] = {0.65, 0.05, 0.05}, green[3] = {0.12, 0.45, 0.15};
    const double lit[3] = {1.0, 1.0, 1.0};
    const double x0 = -12, x1 = 8, y0 = -6, y1 = 6, z0 = -6, z1 = 6;
    {   // back wall x = x1
        double a[3] = {x1, y0, z0}, b[3] = {x1, y1, z0}, c[3] = {x1, y1, z1}, d[3] = {x1, y0, z1};
        add_quad(s, a, b, c, d, white);
    }
    {   // floor z = z0
        double a[3] = {x0, y0, z0}, b[3] = {x1, y0, z0}, c[3] = {x1, y1, z0}, d[3] = {x0, y1, z0};
        add_quad(s, a, b, c, d, white);
    }
    {   // ceiling z = z1
        double a[3] = {x0, y0, z1}, b[3] = {x0, y1, z1}, c[3] = {x1, y1, z1}, d[3] = {x1, y0, z1};
        add_quad(s, a, b, c, d, white);
    }
    {   // left wall y = y1 (camera-left is +y)
        double a[3] = {x0, y1, z0}, b[3] = {x1, y1, z0}, c[3] = {x1, y1, z1}, d[3] = {x0, y1, z1};
        add_quad(s, a, b, c, d, red);
    }
    {   // right wall y = y0
        double a[3] = {x0, y0, z0}, b[3] = {x0, y0, z1}, c[3] = {x1, y0, z1}, d[3] = {x1, y0, z0};
        add_quad(s, a, b, c, d, green);
    }
    {   // area light just under the ceiling
        const double zl = z1 - 0.02;
        double a[3] = {-1.5, -2, zl}, b[3] = {3.5, -2, zl}, c[3] = {3.5, 2, zl}, d[3] = {-1.5, 2, zl};
        add_quad(s, a, b, c, d, lit, G19_BSDF_EMITTER, 17.f);
    }
    add_sphere(s, 3.0, 2.6, -4.0, 2.f, white, glass ? G19_BSDF_MIRROR : G19_BSDF_DIFFUSE);
    add_sphere(s, -0.5, -2.6, -4.0, 2.f, white, glass ? G19_BSDF_GLASS : G19_BSDF_DIFFUSE);
}

void heightfield(g19_scene* s, int n, bool room) {
    const double grey[3] = {0.7, 0.7, 0.7}, lit[3] = {1.0, 1.0, 1.0};
    if (room) {
        // C4 as a CLOSED room (a scene where bounces survive): the surface is the far wall of a box
        // x in [-12, 8], y,z in [-8, 8] with a ceiling light; walls first, light next, the mesh last (far-to-near
        // does not apply: the mesh is the farthest thing the camera sees)
        const double white[3] = {0.73, 0.73, 0.73}, red[3] = {0.65, 0.05, 0.05}, green[3] = {0.12, 0.45, 0.15};
        const double x0 = -12, x1 = 8, y0 = -8, y1 = 8, z0 = -8, z1 = 8;
        { double a[3] = {x1, y0, z0}, b[3] = {x1, y1, z0}, c[3] = {x1, y1, z1}, d[3] = {x1, y0, z1}; add_quad(s, a, b, c, d, white); }
        { double a[3] = {x0, y0, z0}, b[3] = {x0, y0, z1}, c[3] = {x0, y1, z1}, d[3] = {x0, y1, z0}; add_quad(s, a, b, c, d, white); }
        { double a[3] = {x0, y0, z0}, b[3] = {x1, y0, z0}, c[3] = {x1, y1, z0}, d[3] = {x0, y1, z0}; add_quad(s, a, b, c, d, white); }
        { double a[3] = {x0, y0, z1}, b[3] = {x0, y1, z1}, c[3] = {x1, y1, z1}, d[3] = {x1, y0, z1}; add_quad(s, a, b, c, d, white); }
        { double a[3] = {x0, y1, z0}, b[3] = {x1, y1, z0}, c[3] = {x1, y1, z1}, d[3] = {x0, y1, z1}; add_quad(s, a, b, c, d, red); }
        { double a[3] = {x0, y0, z0}, b[3] = {x0, y0, z1}, c[3] = {x1, y0, z1}, d[3] = {x1, y0, z0}; add_quad(s, a, b, c, d, green); }
        const double zl = z1 - 0.02;
        double a[3] = {-5, -4, zl}, b[3] = {3, -4, zl}, c[3] = {3, 4, zl}, d[3] = {-5, 4, zl};
        add_quad(s, a, b, c, d, lit, G19_BSDF_EMITTER, 9.f);
    }
    auto vertex = [n](int j, int k, double* p) {
        double y = -8.0 + 16.0 * double(j) / double(n);
        double z = -8.0 + 16.0 * double(k) / double(n);
        set3(p, 6.0 + 1.5 * std::sin(0.7 * y) * std::cos(0.9 * z), y, z);
    };
    s->ents.reserve(s->ents.size() + size_t(2) * n * n + 2);
    for (int k = 0; k < n; ++k) {
        for (int j = 0; j < n; ++j) {
            double a[3], b[3], c[3], d[3];
            vertex(j, k, a);
            vertex(j + 1, k, b);
            vertex(j + 1, k + 1, c);
            vertex(j, k + 1, d);
            add_quad(s, a, b, c, d, grey);
        }
    }
    if (room) return;
    // a large emitter behind and above the camera, facing the surface
    double a[3] = {-14, -9, -9}, b[3] = {-14, 9, -9}, c[3] = {-14, 9, 9}, d[3] = {-14, -9, 9};
    add_quad(s, a, b, c, d, lit, G19_BSDF_EMITTER, 2.f);
}

} // namespace

int make_builtin(int which, int n, int w, int h, g19_scene** out, g19_camera* cam, double light[3]) {
    if (!out) return G19_ERR_INVALID;
    g19_scene* s = nullptr;
    int rc = new_scene(&s, 20.0);
    if (rc != G19_OK) return rc;
    switch (which) {
    case G19_SCENE_DEFAULT: {
        // main.cpp:24-57: ExpQuad first, then the two ImpSpheres
        g19_entity_desc q = blank(G19_EXP_QUAD);
        set3(q.p, 0, 0, 0);
        q.f[0] = 2;
        q.f[1] = 3;
        q.f[2] = float(90.0 * M_PI / 180.0);
        set3(q.color, 1, 2, 3);
        add(s, q);
        const double redc[3] = {1, 0, 0}, bluec[3] = {0, 0, 1};
        add_sphere(s, 3, 4, 4, 2.f, redc, G19_BSDF_DIFFUSE);
        add_sphere(s, 4, -4, 4, 2.f, bluec, G19_BSDF_DIFFUSE);
        if (cam) {
            set3(cam->pos, -10, 0, 0);
            set3(cam->look_at, 1, 0, 0);
            cam->focal = 0.1;
        }
        if (light) set3(light, -10, 10, 10);
        break;
    }
    case G19_SCENE_CORNELL:
    case G19_SCENE_CORNELL_GLASS:
        cornell(s, which == G19_SCENE_CORNELL_GLASS);
        // the pixel pitch is fixed (raytracer.h:26), so the field of view grows with the
        // width; scale the focal length with it to frame the room the same at any size
        centred_camera(cam, -10, 0, 0, 0.2 * double(w) / 1920.0, w, h);
        if (light) set3(light, 1.0, 0.0, 5.5);
        break;
    case G19_SCENE_HEIGHTFIELD:
    case G19_SCENE_HEIGHTFIELD_ROOM:
        if (n < 1 || n > 4096) {
            g19_scene_destroy(s);
            return G19_ERR_INVALID;
        }
        heightfield(s, n, which == G19_SCENE_HEIGHTFIELD_ROOM);
        centred_camera(cam, -10, 0, 0, 0.2 * double(w) / 1920.0, w, h);
        if (light) set3(light, -10, 10, 10);
        break;
    default: g19_scene_destroy(s); return G19_ERR_INVALID;
    }
    *out = s;
    return G19_OK;
}

} // namespace g19
