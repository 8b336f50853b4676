// material.h -- interface-compatible Material (reference include/material.h:12-107).
// The data members are the reference's; the two shading functions run the
// engine's device code for one point (g19_probe_shade) instead of evaluating
// Blinn-Phong on the host, so a caller sees the same colours the renderer uses.
#pragma once
#include "g19/compat.h"
#include "g19/probe.h"
#include "ray.h"

struct Material {
    explicit Material(glm::dvec3 c) : color(c), diffuse_color(c * 0.5), specular_color(1, 1, 1) {}
    Material(glm::dvec3 c, glm::dvec3 shader) : color(c), diffuse_color(c * 0.5), specular_color(1, 1, 1), shader_parameters(shader) {}

    glm::dvec3 color;
    glm::dvec3 diffuse_color;
    glm::dvec3 specular_color;
    glm::dvec3 shader_parameters = glm::dvec3(0.1, 0.7, 1);
    double specular_power = 5;

    // PATH-mode extension (not in the reference): bounce model of the surface.
    int bsdf = G19_BSDF_DIFFUSE;
    glm::dvec3 emission = glm::dvec3(0, 0, 0);
    double ior = 1.5;

    glm::dvec3 blinn_phong(Ray ray, glm::dvec3 light, glm::dvec3 intersect, glm::dvec3 normal) const {
        return g19::detail::shade_point(color, diffuse_color, specular_color, shader_parameters, specular_power, 0, ray.dir, light,
                                        intersect, normal, 0, 0);
    }
    glm::dvec3 blinn_phong_texture(Ray ray, glm::dvec3 light, glm::dvec3 intersect, glm::dvec3 normal, int relative_x,
                                   int relative_y) const {
        return g19::detail::shade_point(color, diffuse_color, specular_color, shader_parameters, specular_power, 1, ray.dir, light,
                                        intersect, normal, relative_x, relative_y);
    }
};
