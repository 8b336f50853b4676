// camera.h -- interface-compatible Camera (reference include/camera.h:6-17).
// Plain data consumed by RayTracer; the basis (left, top_left) is derived from
// it inside lib2019global_b200 exactly as raytracer.h:26-30 does.
#pragma once
#include "g19/compat.h"

struct Camera {
    explicit Camera(glm::dvec3 position) : Camera(position, glm::dvec3(0, 0, 0), 0.04) {}
    Camera(glm::dvec3 position, glm::dvec3 lookAt, double focal)
        : pos(position), up(0, 0, 1.0), forward(glm::normalize(lookAt - position)), lookAtPoint(lookAt), focalDist(focal) {}

    glm::dvec3 pos;
    glm::dvec3 up;
    glm::dvec3 forward; // unit view direction
    glm::dvec3 lookAtPoint; // (addition) kept so the engine recomputes `forward` with the reference's own ops
    const double sensorDiag = 0.035;
    const double focalDist = 0.04;
};
