// g19/probe.h -- one-entity GPU probes behind the host-callable virtuals of the
// interface classes (Entity::intersect, getTextureCoord, Material::blinn_phong*,
// Octree::intersect). The reference evaluates these on the CPU; here each call
// uploads the entity and runs the SAME device function the render kernels use
// for one ray / one point. They exist for interface completeness and for
// tests; RayTracer::run never goes through them.
#pragma once
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include "g19/compat.h"

namespace g19 {
namespace detail {

inline void check(int rc, g19_ctx* ctx, const char* what) {
    if (rc != G19_OK) throw std::runtime_error(std::string(what) + ": " + g19_last_error(ctx));
}

// Lazily created engine context shared by all probes of the process.
inline g19_ctx* probe_ctx() {
    static g19_ctx* ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int rc = g19_create(nullptr, 0, &ctx);
        if (rc != G19_OK) throw std::runtime_error(std::string("g19_create: ") + g19_last_error(nullptr));
    });
    return ctx;
}

inline std::mutex& probe_mutex() {
    static std::mutex m;
    return m;
}

// What the probe context currently holds: a whole Octree's scene (handle, version) or a single entity's
// descriptor -- repeated probes of the same thing skip the upload (callers hold probe_mutex()).
struct ProbeCache {
    unsigned long serial = 0; // Octree::_serial (process-unique, never reused), 0 = none
    unsigned version = 0;
    bool single = false;
    g19_entity_desc desc;
};
inline ProbeCache& probe_cache() {
    static ProbeCache c;
    return c;
}

inline unsigned long next_serial() {
    static std::mutex m;
    static unsigned long n = 0;
    std::lock_guard<std::mutex> lock(m);
    return ++n;
}

inline void upload_cached(const g19_scene* s, unsigned long serial, unsigned version) {
    ProbeCache& c = probe_cache();
    if (!c.single && c.serial == serial && c.version == version) return;
    c.serial = 0;
    check(g19_upload_scene(probe_ctx(), s), probe_ctx(), "g19_upload_scene");
    c.serial = serial;
    c.version = version;
    c.single = false;
}

// Uploads a scene holding exactly `d` (a root box that always accepts it).
inline void load_single(const g19_entity_desc& d) {
    ProbeCache& c = probe_cache();
    if (c.single && std::memcmp(&c.desc, &d, sizeof d) == 0) return;
    c.single = false;
    c.serial = 0;
    const double lo[3] = {-1e30, -1e30, -1e30}, hi[3] = {1e30, 1e30, 1e30};
    g19_scene* s = nullptr;
    check(g19_scene_create(lo, hi, &s), nullptr, "g19_scene_create");
    int rc = g19_scene_add_entity(s, &d, nullptr);
    if (rc == G19_OK) rc = g19_upload_scene(probe_ctx(), s);
    g19_scene_destroy(s);
    check(rc, probe_ctx(), "g19_scene_add_entity / g19_upload_scene");
    std::memcpy(&c.desc, &d, sizeof d);
    c.single = true;
}

inline bool intersect_one(const g19_entity_desc& d, glm::dvec3 o, glm::dvec3 dir, glm::dvec3& point, glm::dvec3& normal) {
    std::lock_guard<std::mutex> lock(probe_mutex());
    load_single(d);
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {dir.x, dir.y, dir.z};
    int32_t hit = 0;
    double p[3], n[3];
    check(g19_probe_intersect(probe_ctx(), 0, 1, oo, dd, &hit, p, n), probe_ctx(), "g19_probe_intersect");
    if (hit) {
        point = glm::dvec3(p[0], p[1], p[2]);
        normal = glm::dvec3(n[0], n[1], n[2]);
    }
    return hit != 0;
}

inline void texcoord_one(const g19_entity_desc& d, glm::dvec3 at, int& u, int& v) {
    std::lock_guard<std::mutex> lock(probe_mutex());
    load_single(d);
    const double p[3] = {at.x, at.y, at.z};
    int32_t uv[2] = {0, 0};
    check(g19_probe_texcoord(probe_ctx(), 0, 1, p, uv), probe_ctx(), "g19_probe_texcoord");
    u = uv[0];
    v = uv[1];
}

inline glm::dvec3 shade_point(glm::dvec3 color, glm::dvec3 diffuse, glm::dvec3 specular, glm::dvec3 shader, double power,
                              int textured, glm::dvec3 dir, glm::dvec3 light, glm::dvec3 at, glm::dvec3 normal, int u, int v) {
    std::lock_guard<std::mutex> lock(probe_mutex());
    g19_entity_desc d = {};
    d.kind = G19_IMP_SPHERE; // any entity carries the Material
    d.f[0] = 1.f;
    d.color[0] = color.x; d.color[1] = color.y; d.color[2] = color.z;
    d.material_set = 1;
    d.diffuse_color[0] = diffuse.x; d.diffuse_color[1] = diffuse.y; d.diffuse_color[2] = diffuse.z;
    d.specular_color[0] = specular.x; d.specular_color[1] = specular.y; d.specular_color[2] = specular.z;
    d.shader_parameters[0] = shader.x; d.shader_parameters[1] = shader.y; d.shader_parameters[2] = shader.z;
    d.specular_power = power;
    load_single(d);
    const double dd[3] = {dir.x, dir.y, dir.z}, ll[3] = {light.x, light.y, light.z}, pp[3] = {at.x, at.y, at.z},
                 nn[3] = {normal.x, normal.y, normal.z};
    double rgb[3];
    check(g19_probe_shade(probe_ctx(), 0, textured, dd, ll, pp, nn, u, v, rgb), probe_ctx(), "g19_probe_shade");
    return glm::dvec3(rgb[0], rgb[1], rgb[2]);
}

} // namespace detail
} // namespace g19
