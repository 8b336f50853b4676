// device_scene.h -- flat, cache-aligned records the kernels read from HBM.
//
// Two flattened views of the same host scene are uploaded:
//   REF  : the reference's own octree, bug for bug (explicit FP64 node boxes,
//          leaf entity lists with duplicates, entity records that expand to
//          FP64 triangle records). Feeds ref_kernels.cu. Bit-exact contract.
//   PATH : this engine's linear octree over FP32 primitives (implicit node
//          boxes, 8-byte node records, 8 children = one 64 B line). Feeds
//          path_kernels.cu.
#pragma once

#include <cstdint>

namespace g19 {

// ---- REF view ---------------------------------------------------------------
struct alignas(64) RefNodeD { // Octree::Node, octree.h:157-160
    double mn[3], mx[3];
    int32_t first_child; // -1: leaf; children are first_child .. first_child+7
    int32_t ent_offset;  // into ref_ents[] (meaningful for leaves)
    int32_t ent_count;   // Node::_entities.size() -- also for interior nodes (octree.h:140)
    int32_t pad;
};
static_assert(sizeof(RefNodeD) == 64, "one node per 64 B line");

struct alignas(32) RefTriD { // ImpTriangle, entities.h:142-148
    double p1[3], p2[3], p3[3];
    double pos[3], normal[3];
    float e1[3], e2[3]; // edge1/edge2 as glm::mat3's constructor narrows them
    int32_t pad[2];
};
static_assert(sizeof(RefTriD) == 160, "RefTriD layout");

struct alignas(64) RefEntityD {
    int32_t kind, combine;
    int32_t tri_offset, tri_count; // into ref_tris[]
    int32_t first_tested;
    float radius;
    float f[4];        // ctor scalars (width, length, ...)
    double pos[3];
    double color[3];   // Material::color
    double aux0[3], aux1[3];
    // Material's remaining public fields (material.h:23-29)
    double diffuse_color[3], specular_color[3], shader[3];
    double specular_power;
    int32_t pad[4];
};
static_assert(sizeof(RefEntityD) == 256, "RefEntityD layout");

// ---- PATH view --------------------------------------------------------------
// Hot intersection record, 64 B = 4 x float4 (one line per two primitives):
//   triangle: rows 0..2 = the affine map world -> (b1, b2, h): barycentric coordinates of the
//             foot point and the height above the plane in units of the normal, so that a ray
//             o + t d hits at t = -h(o)/h'(d) with (u,v) = (b1,b2)(o) + t (b1,b2)'(d)
//   sphere  : row 0 = (centre.xyz, radius)
//   row 3   = (material index bits, bsdf bits, kind: 1 triangle / 0 sphere, 0)
struct alignas(16) PrimHot {
    float q[16];
};
// Cold shading record, 32 B = 2 x float4: everything a surface interaction needs in ONE hop
// from the hit record -- geometric normal (triangles), index of refraction, albedo, and the
// material index (emitters look their radiance up in materials[]).
struct alignas(16) PrimCold {
    float n[3];
    float ior;
    float albedo[3];
    int32_t material; // index into materials[]
};
struct alignas(16) MaterialD {
    float albedo[3];
    int32_t bsdf;
    float emission[3];
    float ior;
};
// Linear octree node, 8 B. Children of one node are contiguous and complete
// (always 8 records = one 64 B line); empty octants have count == 0.
//   leaf    : first = offset into prim_index[], count = number of primitives,
//             bit 31 of `count` set
//   interior: first = index of child 0, count = 0
struct PathNodeD {
    uint32_t first;
    uint32_t count;
};
constexpr uint32_t kLeafBit = 0x80000000u;

// Area-light table entry (triangle emitters), for next-event estimation.
struct alignas(16) LightD {
    float v0[3], area;
    float e1[3], pdf_pick; // probability of picking this light
    float e2[3];
    int32_t prim;
    float n[3];
    float pad;
    float emission[3];
    float pad2;
};

} // namespace g19
