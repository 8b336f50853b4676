// bbox.h -- interface-compatible BoundingBox (reference include/bbox.h:10-52).
// Host-side, build-time only: the octree build inside
// lib2019global_b200 (csrc/scene.cpp) applies the same strict-< overlap rule.
#pragma once
#include <cassert>
#include <cmath>
#include "g19/compat.h"

struct BoundingBox {
    BoundingBox(glm::dvec3 lo, glm::dvec3 hi) : min(lo), max(hi) {}
    double dx() const { return max.x - min.x; }
    double dy() const { return max.y - min.y; }
    double dz() const { return max.z - min.z; }
    const glm::dvec3 min;
    const glm::dvec3 max;

    bool intersect(const BoundingBox& o) const { // centre distance < sum of half extents, per axis
        for (int k = 0; k < 3; ++k) {
            double gap = 0.5 * (min[k] + max[k]) - 0.5 * (o.min[k] + o.max[k]);
            if (!(std::fabs(gap) < 0.5 * (max[k] - min[k]) + 0.5 * (o.max[k] - o.min[k]))) return false;
        }
        return true;
    }
    bool contains(glm::dvec3 p) const {
        for (int k = 0; k < 3; ++k)
            if (!(std::fabs(0.5 * (min[k] + max[k]) - p[k]) <= 0.5 * (max[k] - min[k]))) return false;
        return true;
    }
};
