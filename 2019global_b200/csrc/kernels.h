// kernels.h -- launch entry points of the two CUDA translation units.
//   ref_kernels.cu  (compiled with -fmad=false: bit-exact reference arithmetic)
//   path_kernels.cu (FP32 wavefront path tracer)
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "device_scene.h"

namespace g19 {

constexpr int kTile = 32;              // image tile edge
constexpr int kTilePix = kTile * kTile;

// Interleaved tile ownership: global tile t belongs to rank t % world. A rank
// addresses its pixels through a compact "local pixel" index
//   lp = local_tile * 1024 + ly * 32 + lx,   global tile = local_tile * world + rank.
struct TileMap {
    int32_t w, h;
    int32_t tiles_x, tiles_y;
    int32_t rank, world;
    int32_t n_local_tiles;
    int32_t n_local_pix; // n_local_tiles * 1024 (includes out-of-frame padding)
};

inline TileMap make_tile_map(int w, int h, int rank, int world) {
    TileMap m;
    m.w = w;
    m.h = h;
    m.tiles_x = (w + kTile - 1) / kTile;
    m.tiles_y = (h + kTile - 1) / kTile;
    m.rank = rank;
    m.world = world;
    int n = m.tiles_x * m.tiles_y;
    m.n_local_tiles = (n > rank) ? (n - rank + world - 1) / world : 0;
    m.n_local_pix = m.n_local_tiles * kTilePix;
    return m;
}

// Camera basis exactly as raytracer.h:26-30 evaluates it (FP64, host side).
struct RefCamera {
    double pos[3], up[3], left[3], top_left[3];
    double light[3];
};

struct RefSceneD {
    const RefNodeD* nodes;
    const int32_t* ents;      // concatenated leaf entity lists
    const RefEntityD* entities;
    const RefTriD* tris;
    int32_t n_nodes, n_entities;
};

// ---- ref_kernels.cu ------------------------------------------------------
// raygen + octree traversal + entity intersection + front-object selection
// (raytracer.h:41-74). Outputs are indexed by local pixel.
// [lp0, lp1) = the band of this rank's local pixels to render (lp1 < 0: all): the host-pointer entry points render
// heavy scenes band by band so that stop() and the viewer's repaint see the frame grow (raytracer.h:32-33).
// Heavy rays (ref_heavy_kernel): a ray whose walk exceeds `budget` node expansions is put on a list and spread over
// many warps, one subtree of level kHeavyLevel each. Device buffers, one set per band in flight.
constexpr int kHeavyLevel = 4;
constexpr int kHeavyTasks = 1 << (3 * kHeavyLevel);       // subtrees per ray
constexpr int kHeavyCache = 1 + 8 + 64 + 512;             // child-box masks of the nodes above the subtrees
struct RefHeavyD {
    unsigned* count;        // [0] rays appended (may run past cap), [1] task counter
    int32_t* lp;            // [cap] local pixel of each heavy ray
    unsigned* best;         // [cap] lowest subtree code with a published hit (0xffffffff: none)
    int* lock;              // [cap]
    unsigned short* masks;  // [cap][kHeavyCache]
    int cap, budget;        // budget 0 = off
};
inline size_t ref_heavy_bytes(int cap) { return 64 + size_t(cap) * (4 + 4 + 4 + 2 * kHeavyCache + 2); }
void launch_ref_visibility(const RefSceneD& scene, const RefCamera& cam, const TileMap& map, int32_t* ids,
                           double* points, double* normals, unsigned long long* counters /*nullable: [node,prim]*/,
                           cudaStream_t stream, int lp0 = 0, int lp1 = -1, unsigned* next = nullptr /* device work counter: enables the warp-per-ray kernel */,
                           const RefHeavyD* heavy = nullptr);
// getTextureCoord + blinn_phong_texture + RGB888 quantisation (raytracer.h:76-82).
void launch_ref_shade(const RefSceneD& scene, const RefCamera& cam, const TileMap& map, const int32_t* ids,
                      const double* points, const double* normals, uint8_t* rgb, float* colour, cudaStream_t stream,
                      int lp0 = 0, int lp1 = -1);
// Probes: Entity::intersect on n rays; Octree::intersect candidate list of one ray.
void launch_probe_intersect(const RefSceneD& scene, int32_t entity, int n, const double* origins, const double* dirs,
                            int32_t* hit, double* points, double* normals, cudaStream_t stream);
void launch_probe_candidates(const RefSceneD& scene, const double* origin_dir6, int32_t* out_ids, int max_out,
                             int32_t* out_n, cudaStream_t stream);

// Entity::getTextureCoord on n points; Material::blinn_phong(_texture) on one point.
void launch_probe_texcoord(const RefSceneD& scene, int32_t entity, int n, const double* points, int32_t* uv,
                           cudaStream_t stream);
void launch_probe_shade(const RefSceneD& scene, int32_t entity, int textured, const double* in15, int u, int v,
                        double* out_rgb, cudaStream_t stream);

// ---- shared (ref_kernels.cu) ----------------------------------------------
// Scatter compact local-pixel buffers into full-frame row-major buffers.
void launch_untile(const TileMap& map, const uint8_t* rgb_local, const int32_t* ids_local, const float* rad_local,
                   uint8_t* rgb_frame, int32_t* ids_frame, float* rad_frame, cudaStream_t stream);


// ---- frame_kernels.cu: resolve fused with the multi-GPU framebuffer gather ----------------
void launch_resolve_to_frame(const TileMap& map, const float* accum, int spp, uint8_t* rgb_frame, float* rad_frame,
                             cudaStream_t stream);
void launch_frame_signal(unsigned* flags, cudaStream_t stream);                  // arrived += 1 (system scope)
// `status` (nullable): a word in this process's mapped host memory that a spin which gives up (5 s) sets to 1
void launch_frame_wait(unsigned* flags, unsigned target, unsigned* status, cudaStream_t stream);   // owner: until arrived >= target
void launch_frame_release(unsigned* flags, unsigned epoch, cudaStream_t stream); // owner: consumed = epoch
void launch_frame_acquire(unsigned* flags, unsigned need, unsigned* status, cudaStream_t stream);  // writer: until consumed >= need

} // namespace g19
