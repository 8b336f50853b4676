// octree.h -- interface-compatible Octree (reference include/octree.h:12-68).
//
// push_back forwards each entity's descriptor to a g19_scene, where the
// reference's build (partition / push_obj, octree.h:75-129) is mirrored bug for
// bug on the host (csrc/scene.cpp) and later flattened to HBM. intersect() is a
// GPU probe returning the reference's candidate list. handle() is the one
// addition the reference cannot offer (its Node/_root are private): the way for
// RayTracer to reach the tree.
#pragma once
#include <memory>
#include <vector>
#include "bbox.h"
#include "entities.h"

class Octree {
  public:
    Octree(glm::dvec3 lo, glm::dvec3 hi) : min(lo), max(hi) {
        const double a[3] = {lo.x, lo.y, lo.z}, b[3] = {hi.x, hi.y, hi.z};
        g19_scene* s = nullptr;
        g19::detail::check(g19_scene_create(a, b, &s), nullptr, "g19_scene_create");
        _scene.reset(s, g19_scene_destroy);
    }

    glm::dvec3 min;
    glm::dvec3 max;

    void push_back(Entity* object) {
        g19_entity_desc d = object->describe();
        // a root-test reject is silent, as in octree.h:22-24 (the engine still numbers the entity); anything else
        // means the engine did NOT number it, and appending it here would shift every later id
        const int rc = g19_scene_add_entity(_scene.get(), &d, nullptr);
        if (rc != G19_OK && rc != G19_ERR_REJECTED) g19::detail::check(rc, nullptr, "g19_scene_add_entity");
        _entities.push_back(object);
        ++_version;
    }

    std::vector<Entity*> intersect(const Ray& ray) const {
        std::lock_guard<std::mutex> lock(g19::detail::probe_mutex());
        g19_ctx* ctx = g19::detail::probe_ctx();
        g19::detail::upload_cached(_scene.get(), _serial, _version);
        const double o[3] = {ray.origin.x, ray.origin.y, ray.origin.z}, d[3] = {ray.dir.x, ray.dir.y, ray.dir.z};
        int n = 0;
        std::vector<int32_t> ids(1 << 16);
        g19::detail::check(g19_probe_candidates(ctx, o, d, ids.data(), int(ids.size()), &n), ctx, "g19_probe_candidates");
        std::vector<Entity*> out;
        for (int i = 0; i < n && i < int(ids.size()); ++i) out.push_back(_entities[size_t(ids[i])]);
        return out;
    }

    // additions
    const g19_scene* handle() const { return _scene.get(); }
    unsigned version() const { return _version; }
    const std::vector<Entity*>& entities() const { return _entities; }

  private:
    std::shared_ptr<g19_scene> _scene;
    std::vector<Entity*> _entities; // id = push order, same numbering as the engine
    unsigned _version = 0;
    unsigned long _serial = g19::detail::next_serial(); // identifies this Octree to the probe upload cache
};
