/* oracle_scenes.c -- the procedural scenes of BASELINE.json `configs`, restated for the CPU checkers.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_scene.h). bench.py's reference arm and cpu_baseline leg build their scenes
 * HERE, so that timing the CPU path never loads the product library; tests/test_path_oracle.py checks that every
 * descriptor equals the one the product's generator (2019global_b200/csrc/builtin_scenes.cpp) emits.
 *
 * Reference counterparts: config 1 is the literal of reference main.cpp:24-57; the other scenes are this repo's
 *  and are built from the reference's own entity classes (ImpTriangle entities.h:136-306,
 * ImpSphere entities.h:43-133) in push order. The camera of the non-square configs is pitched so that the
 * reference's "width used for the vertical extent" (raytracer.h:30) is cancelled (SURVEY.md hard part 4).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_scene.h"

typedef struct {
    g19_entity_desc* v;
    int n, cap;
} dlist;

static void push(dlist* l, const g19_entity_desc* d) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 64;
        l->v = realloc(l->v, sizeof(g19_entity_desc) * (size_t)l->cap);
    }
    l->v[l->n++] = *d;
}

static g19_entity_desc blank(int kind) {
    g19_entity_desc d;
    memset(&d, 0, sizeof d);
    d.kind = kind;
    d.bsdf = G19_BSDF_DIFFUSE;
    d.color[0] = 1.0; /* Entity() default material, entities.h:21 */
    d.ior = 1.5f;
    return d;
}

static void tri(dlist* l, const double a[3], const double b[3], const double c[3], const double col[3], int bsdf, float emit) {
    g19_entity_desc d = blank(G19_IMP_TRIANGLE);
    memcpy(d.p, a, 24); memcpy(d.p + 3, b, 24); memcpy(d.p + 6, c, 24);
    memcpy(d.color, col, 24);
    d.bsdf = bsdf;
    d.emission[0] = d.emission[1] = d.emission[2] = emit;
    push(l, &d);
}
/* a planar quad a-b-c-d as the triangles (a,b,c) and (a,c,d) */
static void quad(dlist* l, const double a[3], const double b[3], const double c[3], const double d[3], const double col[3],
                 int bsdf, float emit) {
    tri(l, a, b, c, col, bsdf, emit);
    tri(l, a, c, d, col, bsdf, emit);
}
static void sphere(dlist* l, double x, double y, double z, float r, const double col[3], int bsdf) {
    g19_entity_desc d = blank(G19_IMP_SPHERE);
    d.p[0] = x; d.p[1] = y; d.p[2] = z;
    d.f[0] = r;
    memcpy(d.color, col, 24);
    d.bsdf = bsdf;
    push(l, &d);
}
#define Q(l, col, bsdf, emit, ax, ay, az, bx, by, bz, cx, cy, cz, dx, dy, dz)                         \
    do {                                                                                             \
        double a_[3] = {ax, ay, az}, b_[3] = {bx, by, bz}, c_[3] = {cx, cy, cz}, d_[3] = {dx, dy, dz}; \
        quad(l, a_, b_, c_, d_, col, bsdf, emit);                                                    \
    } while (0)

static void cornell(dlist* l, int glass) {
    const double white[3] = {0.73, 0.73, 0.73}, red[3] = {0.65, 0.05, 0.05}, green[3] = {0.12, 0.45, 0.15}, lit[3] = {1, 1, 1};
    const double x0 = -12, x1 = 8, y0 = -6, y1 = 6, z0 = -6, z1 = 6, zl = 6 - 0.02;
    Q(l, white, 0, 0.f, x1, y0, z0, x1, y1, z0, x1, y1, z1, x1, y0, z1); /* back wall */
    Q(l, white, 0, 0.f, x0, y0, z0, x1, y0, z0, x1, y1, z0, x0, y1, z0); /* floor */
    Q(l, white, 0, 0.f, x0, y0, z1, x0, y1, z1, x1, y1, z1, x1, y0, z1); /* ceiling */
    Q(l, red, 0, 0.f, x0, y1, z0, x1, y1, z0, x1, y1, z1, x0, y1, z1);   /* left wall (camera-left is +y) */
    Q(l, green, 0, 0.f, x0, y0, z0, x0, y0, z1, x1, y0, z1, x1, y0, z0); /* right wall */
    Q(l, lit, G19_BSDF_EMITTER, 17.f, -1.5, -2, zl, 3.5, -2, zl, 3.5, 2, zl, -1.5, 2, zl);
    sphere(l, 3.0, 2.6, -4.0, 2.f, white, glass ? G19_BSDF_MIRROR : G19_BSDF_DIFFUSE);
    sphere(l, -0.5, -2.6, -4.0, 2.f, white, glass ? G19_BSDF_GLASS : G19_BSDF_DIFFUSE);
}

static void hf_vertex(int n, int j, int k, double* p) {
    double y = -8.0 + 16.0 * (double)j / (double)n, z = -8.0 + 16.0 * (double)k / (double)n;
    p[0] = 6.0 + 1.5 * sin(0.7 * y) * cos(0.9 * z);
    p[1] = y;
    p[2] = z;
}

static void heightfield(dlist* l, int n, int room) {
    const double grey[3] = {0.7, 0.7, 0.7}, lit[3] = {1, 1, 1};
    if (room) {
        const double white[3] = {0.73, 0.73, 0.73}, red[3] = {0.65, 0.05, 0.05}, green[3] = {0.12, 0.45, 0.15};
        const double x0 = -12, x1 = 8, y0 = -8, y1 = 8, z0 = -8, z1 = 8, zl = 8 - 0.02;
        Q(l, white, 0, 0.f, x1, y0, z0, x1, y1, z0, x1, y1, z1, x1, y0, z1);
        Q(l, white, 0, 0.f, x0, y0, z0, x0, y0, z1, x0, y1, z1, x0, y1, z0);
        Q(l, white, 0, 0.f, x0, y0, z0, x1, y0, z0, x1, y1, z0, x0, y1, z0);
        Q(l, white, 0, 0.f, x0, y0, z1, x0, y1, z1, x1, y1, z1, x1, y0, z1);
        Q(l, red, 0, 0.f, x0, y1, z0, x1, y1, z0, x1, y1, z1, x0, y1, z1);
        Q(l, green, 0, 0.f, x0, y0, z0, x0, y0, z1, x1, y0, z1, x1, y0, z0);
        Q(l, lit, G19_BSDF_EMITTER, 9.f, -5, -4, zl, 3, -4, zl, 3, 4, zl, -5, 4, zl);
    }
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j) {
            double a[3], b[3], c[3], d[3];
            hf_vertex(n, j, k, a);
            hf_vertex(n, j + 1, k, b);
            hf_vertex(n, j + 1, k + 1, c);
            hf_vertex(n, j, k + 1, d);
            quad(l, a, b, c, d, grey, G19_BSDF_DIFFUSE, 0.f);
        }
    if (room) return;
    Q(l, lit, G19_BSDF_EMITTER, 2.f, -14, -9, -9, -14, 9, -9, -14, 9, 9, -14, -9, 9);
}

static void centred_camera(g19_camera* cam, double px, double py, double pz, double focal, int w, int h) {
    double s = ((double)w - (double)h) * 0.5 * 0.0002 / focal;
    if (s > 0.95) s = 0.95;
    if (s < -0.95) s = -0.95;
    double c = sqrt(1.0 - s * s);
    cam->pos[0] = px; cam->pos[1] = py; cam->pos[2] = pz;
    cam->look_at[0] = px + c; cam->look_at[1] = py; cam->look_at[2] = pz - s;
    cam->focal = focal;
}

/* Descriptors of builtin scene `which` (enum g19_builtin_scene) in push order; the root box is +-20 for all of
 * them. Returns the count (< 0: unknown scene); *out is malloc'ed, release with g19o_free. */
int g19o_builtin_descs(int which, int n, int w, int h, g19_entity_desc** out, g19_camera* cam, double light[3]) {
    dlist l = {0, 0, 0};
    g19_camera c;
    double lt[3] = {0, 0, 0};
    memset(&c, 0, sizeof c);
    switch (which) {
    case G19_SCENE_DEFAULT: {
        g19_entity_desc q = blank(G19_EXP_QUAD);
        q.f[0] = 2; q.f[1] = 3; q.f[2] = (float)(90.0 * 3.14159265358979323846 / 180.0);
        q.color[0] = 1; q.color[1] = 2; q.color[2] = 3;
        push(&l, &q);
        const double redc[3] = {1, 0, 0}, bluec[3] = {0, 0, 1};
        sphere(&l, 3, 4, 4, 2.f, redc, G19_BSDF_DIFFUSE);
        sphere(&l, 4, -4, 4, 2.f, bluec, G19_BSDF_DIFFUSE);
        c.pos[0] = -10; c.look_at[0] = 1; c.focal = 0.1;
        lt[0] = -10; lt[1] = 10; lt[2] = 10;
        break;
    }
    case G19_SCENE_CORNELL:
    case G19_SCENE_CORNELL_GLASS:
        cornell(&l, which == G19_SCENE_CORNELL_GLASS);
        centred_camera(&c, -10, 0, 0, 0.2 * (double)w / 1920.0, w, h);
        lt[0] = 1.0; lt[1] = 0.0; lt[2] = 5.5;
        break;
    case G19_SCENE_HEIGHTFIELD:
    case G19_SCENE_HEIGHTFIELD_ROOM:
        if (n < 1 || n > 4096) return -1;
        heightfield(&l, n, which == G19_SCENE_HEIGHTFIELD_ROOM);
        centred_camera(&c, -10, 0, 0, 0.2 * (double)w / 1920.0, w, h);
        lt[0] = -10; lt[1] = 10; lt[2] = 10;
        break;
    default: return -1;
    }
    if (cam) *cam = c;
    if (light) memcpy(light, lt, sizeof lt);
    *out = l.v;
    return l.n;
}

void g19o_free(void* p) { free(p); }

int g19o_scene_add(void* h, const g19_entity_desc* d);
int g19o_scene_add_many(void* h, const g19_entity_desc* d, int count) {
    for (int i = 0; i < count; ++i)
        if (g19o_scene_add(h, d + i) < 0) return -1;
    return count;
}
