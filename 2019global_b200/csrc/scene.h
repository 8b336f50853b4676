// scene.h -- host-side scene model of lib2019global_b200.
//
// Mirrors what the reference keeps on the host: the entities a caller
// constructed (reference include/entities.h) and the Octree they were pushed
// into (include/octree.h). The per-ray work is NOT here -- it lives in the CUDA
// kernels (ref_kernels.cu, path_kernels.cu); this file only prepares the data
// they consume. Construction arithmetic follows the reference's float/double
// mix exactly, because vertex bits decide hit booleans downstream.
#pragma once

#include <cstdint>
#include <vector>

#include "g19.h"

namespace g19 {

struct V3 {
    double x, y, z;
};

// An ImpTriangle with the members the reference derives once at construction
// (entities.h:138-148) -- recomputing them on upload keeps the bits.
struct HostTri {
    V3 p1, p2, p3;
    V3 pos, edge1, edge2, normal;
};

// How an entity combines its triangles into one answer (REF mode).
enum Combine : int32_t {
    COMBINE_SPHERE = 0,  // analytic ImpSphere, no triangles
    COMBINE_SINGLE = 1,  // ImpTriangle
    COMBINE_FIRST = 2,   // ExpRectangle: t1, else t2                 entities.h:326-336
    COMBINE_BOX = 3,     // ExpBox: per face FIRST; last hit face wins entities.h:415-440
    COMBINE_NEAREST = 4  // Quad/Cube/Cone/ExpSphere: min dist, "<="    entities.h:596-620
};

struct HostEntity {
    g19_entity_desc desc;
    int32_t kind;
    int32_t combine;
    int32_t first_tested; // ExpSphere's loop starts at triangle 1 (entities.h:520)
    V3 pos;               // Entity::pos after the constructor body
    float radius;
    V3 bbmin, bbmax;      // Entity::boundingBox()
    V3 aux0, aux1;        // getTextureCoord helpers (vertices(0)/(1), p3/p4, ...)
    std::vector<HostTri> tris;
    bool in_tree;         // false: rejected by Octree::push_back's root test
};

// Octree::Node (octree.h:71-161), index-linked. Children are allocated eight
// at a time, contiguously, in the reference's child order 0..7 (octree.h:94-108).
struct HostNode {
    V3 mn, mx;
    int32_t first_child; // -1: leaf
    std::vector<int32_t> ents;
};

} // namespace g19

struct g19_scene {
    int shapes = 0; // enum g19_shapes: which triangles PATH mode extracts from the composite entities
    g19::V3 rmin, rmax;
    std::vector<g19::HostEntity> ents;
    std::vector<g19::HostNode> nodes; // nodes[0] = root
};

namespace g19 {

// Entity constructors (returns false on an unknown kind).
bool build_entity(const g19_entity_desc& d, HostEntity& out);
// Octree::push_back (octree.h:20-30). Returns false when rejected by the root test.
bool push_back(g19_scene& s, int32_t entity);
int max_depth(const g19_scene& s);

// G19_SHAPES_FIXED (fixed_shapes.cpp): the triangles the entity's constructor meant to build; false = the
// reference's own triangles are right (or the entity has none) and PATH mode keeps them.
bool fixed_triangles(const g19_entity_desc& d, std::vector<HostTri>& out);

// Procedural scenes of BASELINE.json `configs`.
int make_builtin(int which, int n, int w, int h, g19_scene** out, g19_camera* cam, double light[3]);

} // namespace g19
