// entities.h -- interface-compatible Entity hierarchy (reference include/entities.h:19-967).
//
// Same class names, constructor signatures and public data members as the
// reference, so main.cpp's scene literal compiles unchanged. What differs is
// where the work happens: a constructor only records its arguments in a
// g19_entity_desc and asks lib2019global_b200 for the derived public members
// (bounding box, vertices/triangles -- computed with the reference's own
// float/double mix, csrc/scene.cpp); the per-ray virtuals are thin GPU probes
// (g19/probe.h). The hot path never calls a virtual: Octree::push_back hands
// the descriptor to the engine and RayTracer::run renders from the flattened
// copy in HBM.
#pragma once
#include <tuple>
#include <vector>
#include "bbox.h"
#include "g19/compat.h"
#include "g19/probe.h"
#include "material.h"
#include "ray.h"

#ifndef PI
#define PI 3.1415926535
#endif

struct Entity {
    Entity() : material(Material(glm::dvec3(1, 0, 0))) {}
    explicit Entity(const Material& m) : material(m) {}
    virtual ~Entity() {}

    virtual bool intersect(const Ray& ray, glm::dvec3& intersect, glm::dvec3& normal) const {
        return g19::detail::intersect_one(describe(), ray.origin, ray.dir, intersect, normal);
    }
    virtual BoundingBox boundingBox() const {
        double b[6];
        g19_entity_desc d = describe();
        g19_entity_bbox(&d, b);
        return BoundingBox(glm::dvec3(b[0], b[1], b[2]), glm::dvec3(b[3], b[4], b[5]));
    }
    virtual std::tuple<int, int> getTextureCoord(glm::dvec3 intersect) const {
        int u = 0, v = 0;
        g19::detail::texcoord_one(describe(), intersect, u, v);
        return std::make_tuple(u, v);
    }
    virtual std::vector<Entity*>* get_childs() const { return nullptr; }

    // (addition) the constructor arguments + material, as the engine consumes them
    g19_entity_desc describe() const {
        g19_entity_desc d = shape();
        d.color[0] = material.color.x; d.color[1] = material.color.y; d.color[2] = material.color.z;
        d.bsdf = material.bsdf;
        d.emission[0] = float(material.emission.x); d.emission[1] = float(material.emission.y);
        d.emission[2] = float(material.emission.z);
        d.ior = float(material.ior);
        // the remaining public Material fields (reference material.h:24-29), as the caller left them
        d.material_set = 1;
        d.diffuse_color[0] = material.diffuse_color.x; d.diffuse_color[1] = material.diffuse_color.y; d.diffuse_color[2] = material.diffuse_color.z;
        d.specular_color[0] = material.specular_color.x; d.specular_color[1] = material.specular_color.y; d.specular_color[2] = material.specular_color.z;
        d.shader_parameters[0] = material.shader_parameters.x; d.shader_parameters[1] = material.shader_parameters.y;
        d.shader_parameters[2] = material.shader_parameters.z;
        d.specular_power = material.specular_power;
        return d;
    }

    glm::dvec3 pos = glm::dvec3(0, 0, 0);
    Material material;

  protected:
    virtual g19_entity_desc shape() const = 0;
    static void put(double* dst, glm::dvec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
};

class ImpTriangle : public Entity {
  public:
    ImpTriangle(glm::dvec3 a, glm::dvec3 b, glm::dvec3 c) : Entity(), p1(a), p2(b), p3(c) {
        edge1 = p2 - p1;
        edge2 = p3 - p1;
        normal = glm::normalize(glm::cross(edge1, edge2));
        pos = 0.5 * (0.5 * (p1 + p2) + p3);
    }
    glm::dvec3 p1, p2, p3;
    glm::dvec3 edge1, edge2, normal;

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_IMP_TRIANGLE;
        put(d.p, p1); put(d.p + 3, p2); put(d.p + 6, p3);
        return d;
    }
};

namespace g19 {
namespace detail {
// Fills the `vertices` / `triangles` members a composite exposes (entities.h:510-512 etc.).
inline void expand(const g19_entity_desc& d, std::vector<glm::dvec3>* vertices, std::vector<Entity*>* triangles) {
    int n = g19_entity_triangles(&d, nullptr, 0);
    std::vector<double> v(size_t(n) * 9);
    g19_entity_triangles(&d, v.data(), n);
    for (int i = 0; i < n; ++i) {
        glm::dvec3 a(v[9 * i], v[9 * i + 1], v[9 * i + 2]), b(v[9 * i + 3], v[9 * i + 4], v[9 * i + 5]),
            c(v[9 * i + 6], v[9 * i + 7], v[9 * i + 8]);
        if (vertices) { vertices->push_back(a); vertices->push_back(b); vertices->push_back(c); }
        if (triangles) triangles->push_back(new ImpTriangle(a, b, c));
    }
}
} // namespace detail
} // namespace g19

class ImpSphere : public Entity {
  public:
    ImpSphere(glm::dvec3 centre, float r, glm::dvec3 color) : Entity(Material(color)), radius(r) { pos = centre; }
    float radius;

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_IMP_SPHERE;
        put(d.p, pos);
        d.f[0] = radius;
        return d;
    }
};

class ExpRectangle : public Entity {
  public:
    ExpRectangle(glm::dvec3 a, glm::dvec3 b, glm::dvec3 c)
        : Entity(), p1(a), p2(b), p3(c), p4(glm::dvec3(0, 0, 0) + (glm::dvec3(0, 0, 0) - c)), t1(a, b, c), t2(a, b, p4) {
        pos = 0.5 * (p1 + p2);
        normal = glm::normalize(glm::cross(p1 - p3, p2 - p3));
    }
    glm::dvec3 p1, p2, p3; // p1, p2 diagonal
    glm::dvec3 p4;         // as the reference evaluates it: -p3
    glm::dvec3 normal;
    ImpTriangle t1, t2;
    std::vector<Entity*>* get_childs() const override {
        static std::vector<Entity*> copy; // the reference also hands out one process-wide list (entities.h:371)
        if (copy.empty()) { copy.push_back(new ImpTriangle(p1, p2, p3)); copy.push_back(new ImpTriangle(p1, p2, p4)); }
        return &copy;
    }

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_EXP_RECTANGLE;
        put(d.p, p1); put(d.p + 3, p2); put(d.p + 6, p3);
        return d;
    }
};

class ExpBox : public Entity {
  public:
    ExpBox(glm::dvec3 lo, glm::dvec3 hi) : Entity(), min(lo), max(hi) {}
    const glm::dvec3 min;
    const glm::dvec3 max;

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_EXP_BOX;
        put(d.p, min); put(d.p + 3, max);
        return d;
    }
};

class ExpSphere : public Entity {
  public:
    ExpSphere(glm::dvec3 centre, float r, glm::dvec3 color) : Entity(Material(color)), radius(r) {
        pos = centre;
        g19::detail::expand(describe(), &vertices, &triangles);
    }
    float radius;
    int sectornum = 10;
    int stacknum = 10;
    std::vector<glm::dvec3> vertices; // (here: the corner points of `triangles`, three per triangle)
    std::vector<glm::dvec3> normal;
    std::vector<Entity*> triangles;

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_EXP_SPHERE;
        put(d.p, pos);
        d.f[0] = radius;
        return d;
    }
};

class ExpQuad : public Entity {
  public:
    ExpQuad(glm::dvec3 centre, float w, float l, float a, glm::dvec3 color)
        : Entity(Material(color)), width(w), length(l), alpha(a) {
        pos = centre;
        g19::detail::expand(describe(), &vertices, &triangles);
    }
    std::vector<glm::dvec3> vertices;
    std::vector<glm::dvec3> normal;
    std::vector<Entity*> triangles;
    float width, length, alpha;
    std::vector<Entity*>* get_childs() const override { return const_cast<std::vector<Entity*>*>(&triangles); }

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_EXP_QUAD;
        put(d.p, pos);
        d.f[0] = width; d.f[1] = length; d.f[2] = alpha;
        return d;
    }
};

class ExpCube : public Entity {
  public:
    ExpCube(glm::dvec3 centre, float w, float l, float h, glm::dvec3 color)
        : Entity(Material(color)), width(w), length(l), height(h) {
        pos = centre;
        dir = glm::normalize(glm::dvec3(-1, -1, -1));
        g19::detail::expand(describe(), &vertices, &triangles);
    }
    std::vector<glm::dvec3> vertices;
    glm::dvec3 dir;
    glm::dvec3 rotate;
    std::vector<Entity*> triangles;
    float width, length, height;
    std::vector<Entity*>* get_childs() const override { return const_cast<std::vector<Entity*>*>(&triangles); }

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_EXP_CUBE;
        put(d.p, pos);
        d.f[0] = width; d.f[1] = length; d.f[2] = height;
        return d;
    }
};

class ExpCone : public Entity {
  public:
    ExpCone(glm::dvec3 apex, glm::dvec3 axis, float h, float r, glm::dvec3 color)
        : Entity(Material(color)), dir(axis), height(h), radius(r) {
        pos = apex;
        g19::detail::expand(describe(), &vertices, &triangles);
    }
    std::vector<glm::dvec3> vertices;
    std::vector<Entity*> triangles;
    glm::dvec3 dir; // kept as passed; the reference builds the cone along normalize({-1,0,-10}) regardless
    float height, radius;
    std::vector<Entity*>* get_childs() const override { return const_cast<std::vector<Entity*>*>(&triangles); }

  protected:
    g19_entity_desc shape() const override {
        g19_entity_desc d = {};
        d.kind = G19_EXP_CONE;
        put(d.p, pos); put(d.p + 3, dir);
        d.f[0] = height; d.f[1] = radius;
        return d;
    }
};
