// ray.h -- interface-compatible Ray (reference include/ray.h:5-9): the
// constructor normalises the direction.
#pragma once
#include "g19/compat.h"

struct Ray {
    Ray(glm::dvec3 from, glm::dvec3 towards) : origin(from), dir(glm::normalize(towards)) {}
    glm::dvec3 origin;
    glm::dvec3 dir;
};
