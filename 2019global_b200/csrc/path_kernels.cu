// path_kernels.cu -- the wavefront path tracer on sm_100a (G19_MODE_PATH).
//
// One pass traces P = spp_pass * window camera paths, bounce by bounce; up to four passes are in
// flight on their own streams (path.cu). Two families of kernels share the shading code:
//
// FLAT scenes (the whole scene is one leaf of <= 192 primitives, staged in shared memory):
//   raygen_extend_flat   camera ray from the pixel index, nearest hit, vertex record written at the
//                        position the warp's ballot compaction reserved in the material's queue
//   bounce_flat<KIND>    one material queue of one bounce (diffuse / mirror / glass; scenes with
//                        specular materials run the three queues of a bounce in ONE launch,
//                        bounce_flat_all): next-event estimation, BSDF sample, the shadow ray and
//                        the continuation ray traced in one loop over primitive PAIRS (packed
//                        FP32), the new vertex record written into the next bounce's queue. The
//                        queues hold the records themselves (four float4 planes), the path's
//                        radiance rides in the record and is stored once when the path ends.
// TREE scenes (linear octree in HBM / L2):
//   raygen_extend        camera ray + ordered octree walk (all 32 lanes of a warp in fixed-trip rounds)
//   bounce<KIND>         shades a queue of slot indices and queues the vertex's rays
//   trace                persistent walk over the bounce's ray queue with dynamic ray fetch
// Both:
//   accumulate           per pixel, in sample order, radiance -> accumulation buffer
//   resolve              mean, clamp, truncating RGB888 store (Image::setPixel, reference include/image.h:14-16)
//
// There is no separate extend launch after the camera segment: a queue entry always produces
// exactly one continuation ray, so compacting between "shade" and "extend" would only re-read what
// the thread already holds in registers. Every launch is a persistent grid (SM count x resident
// CTAs) that reads its queue length from device memory and is chained to its predecessor by
// programmatic dependent launch, so a frame is enqueued without a single host round trip. The RNG
// is Philox4x32-7 keyed on (global pixel, sample, bounce, stream): results do not depend on
// queue order, pass size, passes in flight or how tiles are split across GPUs.
//
// What the ncu captures drove (profiles/r01_tuning_log.md has the numbers):
//   session 1  * the kernels are issue-bound, not HBM-bound -> triangle test by a precomputed affine
//                map (6 dot products, one MUFU reciprocal, no branches), coplanar triangle pairs
//                merged into parallelograms by the builder
//              * barrier stalls from block-level compaction -> queue space is reserved in
//                warp-private 64-entry chunks: one atomic per 64 outputs, no __syncthreads
//              * long-scoreboard stalls on queue -> state dependent loads -> cp.async prefetch of
//                the next record into shared memory
//   session 2  * shade + extend fused; (pixel, sample) ride in the spare words of the vertex record
//              * flat scenes: the shadow ray and the continuation ray leave the same point, so ONE
//                loop over the primitives tests both (shared origin half, shared loads)
//              * tree scenes: while-while walk in converged rounds, dynamic ray fetch
//   session 3  * slot-indexed state cost the same time per launch whatever the number of live
//                paths (sector waste) -> dense vertex records in the queues
//              * 40 % of the instructions FP32 at 70 % issue utilisation -> primitive pairs through
//                FFMA2 / FMUL2 / FADD2
//              * every kernel ends in a tail -> several passes in flight
//
// Scene access: flat scenes stage everything (records, pairs, shading records, lights) into shared
// memory at kernel start with cp.async.bulk (TMA bulk copy, one mbarrier); tree scenes stage the
// breadth-first prefix of the node array and read the rest through L2 with read-only loads.
#include "path.h"

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace g19 {
namespace {

constexpr int kThreads = 256;
constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kInvalid = 0xffffffffu; // padding entry of a queue / "no primitive"
constexpr uint32_t kChunk = 64;            // queue entries reserved per atomic (>= 32)
constexpr float kRayEps = 1.0e-3f;         // origin offset along the normal (scene units)
constexpr float kPi = 3.14159265358979323846f;
constexpr uint32_t kRayShadow = 0x80000000u;   // ray record flags riding on the slot index (< 2^30)
constexpr uint32_t kRaySpecular = 0x40000000u; // the continuation ray leaves a specular vertex
constexpr uint32_t kRaySlotMask = 0x3fffffffu;

__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// Streaming (evict-first) 16-byte accesses for data that is written once and read once far later (ray queue records).
// Through INTEGER registers: the records carry slot indices and flags in the bit patterns of their w words, which look
// like denormals -- the float4 overloads of __ldcs / __stcs are free to flush them (measured: every shadow ray then
// delivered to slot 0).
__device__ __forceinline__ void st_stream(float4* p, float x, float y, float z, uint32_t w_bits) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(__float_as_uint(x), __float_as_uint(y), __float_as_uint(z), w_bits));
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}

// ---- Philox4x32-7 (Salmon et al. 2011: the fewest rounds that pass BigCrush; three rounds = 36 instructions per vertex
// fewer than the -10 variant of round 1); identical integer stream in oracle/path_oracle.c
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0) {
    uint32_t k1 = 0x32303139u; // "2019"
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t v) { return float(v >> 8) * (1.0f / 16777216.0f); }

// ---- small float3 kit -------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 normalize(float3 v) { return v * rsqrtf(dot(v, v)); }

// ---- shared-memory scene prefix ----------------------------------------------
// Dynamic shared memory, sized on the host to what the scene needs (PassArgs::stage_*):
//   [nodes: stage_nodes x 8 B][hot primitive records: stage_prims x 64 B (flat scenes only)]
//   [cold records: stage_cold x 32 B][lights: stage_lights x 80 B]   (flat scenes only)
//   [traversal stack: stack_levels x kThreads x 4 B][mbarrier: 8 B]
// A Cornell box stages whole (about 0.8 KB) and leaves the SM free for more CTAs.
extern __shared__ __align__(16) unsigned char g19_dyn_smem[];

template <bool ALL> struct SceneAccess {
    const PathSceneD* g;
    const uint2* nodes_s;
    const float4* hot_s;
    const float4* cold_s;
    const float4* lights_s;
    const ulonglong2* pairs_s; // flat scenes: primitive PAIRS, one float2 per coefficient (trace_flat)
    uint32_t* stack;
    int n_nodes_s;
    int walk_steps;      // tree walk: cell moves per round
    uint32_t leaf_batch; // tree walk: primitives tested per round
    int bvh_stack;       // BvhWalk: levels of the postponed-children stack in shared memory (= PassArgs::stack_levels)
    int bvh_spec;        // BvhWalk: a lane that reaches a leaf postpones it and keeps descending (speculative traversal)
    unsigned long long* coop; // tree walk, cooperative leaf tests: this warp's 32 result slots, or nullptr
    __device__ __forceinline__ uint2 node(uint32_t i) const {
        if (ALL || i < (uint32_t)n_nodes_s) return nodes_s[i];
        return __ldg(reinterpret_cast<const uint2*>(g->nodes) + i);
    }
    // shading record of primitive id `prim`: (normal, ior), (albedo, material)
    __device__ __forceinline__ void cold(uint32_t prim, float4& c0, float4& c1) const {
        if (ALL) {
            c0 = cold_s[2 * prim]; c1 = cold_s[2 * prim + 1];
        } else {
            const float4* p = reinterpret_cast<const float4*>(g->cold) + 2 * (size_t)prim;
            c0 = __ldg(p); c1 = __ldg(p + 1);
        }
    }
    // area-light record = 5 x float4: (v0, area) (e1, pdf_pick) (e2, -) (n, -) (emission, -)
    __device__ __forceinline__ const float4* light(int li) const {
        if (ALL) return lights_s + 5 * li;
        return reinterpret_cast<const float4*>(g->lights) + 5 * (size_t)li;
    }
    // row 0 / row 3 of the hot record BY PRIMITIVE ID (flat scenes: leaf order == id order)
    __device__ __forceinline__ float4 hot_row(uint32_t prim, int row) const {
        if (ALL) return hot_s[4 * prim + row];
        return __ldg(reinterpret_cast<const float4*>(g->hot) + 4 * (size_t)prim + row);
    }
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Programmatic dependent launch: a pass is a chain of 7 to 40 short persistent kernels, each
// depending on the previous one's queues. Every kernel lets its successor be scheduled as soon as
// CTAs retire (launch_dependents, first instruction), does whatever does not depend on earlier
// kernels (staging the scene into shared memory), and only then waits for the predecessor's
// memory (griddepcontrol.wait): launch latency and prologue hide under the predecessor's tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Stage the scene prefix with TMA bulk copies (cp.async.bulk) completing on one mbarrier.
template <bool ALL> __device__ __forceinline__ SceneAccess<ALL> stage_scene(const PassArgs& a) {
    const PathSceneD& g = a.scene;
    SceneAccess<ALL> acc;
    acc.g = &a.scene;
    acc.n_nodes_s = a.stage_nodes;
    acc.walk_steps = a.walk_steps;
    acc.leaf_batch = uint32_t(a.leaf_batch);
    acc.bvh_spec = a.bvh_spec;
    acc.bvh_stack = a.stack_levels;
    acc.coop = nullptr;
    // byte counts rounded up to 16 (the device arrays are padded, see DeviceArray::ensure)
    const uint32_t nb = (uint32_t(a.stage_nodes) * 8u + 15u) & ~15u;
    const uint32_t pb = uint32_t(a.stage_prims) * 64u;
    const uint32_t cb = ALL ? uint32_t(a.stage_cold) * 32u : 0u;
    const uint32_t lb = ALL ? uint32_t(a.stage_lights) * 80u : 0u;
    const uint32_t qb = ALL ? uint32_t(g.pairs_bytes) : 0u;
    unsigned char* base = g19_dyn_smem;
    acc.nodes_s = reinterpret_cast<const uint2*>(base);
    acc.hot_s = reinterpret_cast<const float4*>(base + nb);
    acc.cold_s = reinterpret_cast<const float4*>(base + nb + pb);
    acc.lights_s = reinterpret_cast<const float4*>(base + nb + pb + cb);
    acc.pairs_s = reinterpret_cast<const ulonglong2*>(base + nb + pb + cb + lb);
    acc.stack = reinterpret_cast<uint32_t*>(base + nb + pb + cb + lb + qb) + threadIdx.x;
    unsigned long long* barp =
        reinterpret_cast<unsigned long long*>(base + nb + pb + cb + lb + qb + uint32_t(a.stack_levels) * kThreads * 4u);
    const uint32_t bar = smem_addr(barp);
    if (!ALL && a.coop_leaf) acc.coop = barp + 2 + (threadIdx.x >> 5) * 32; // [barrier 8 B, pad 8 B][8 warps x 32 slots x 8 B]
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb + pb + cb + lb + qb) : "memory");
        if (nb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base)),
                         "l"(g.nodes), "r"(nb), "r"(bar)
                         : "memory");
        if (pb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base + nb)),
                         "l"(g.hot), "r"(pb), "r"(bar)
                         : "memory");
        if (cb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base + nb + pb)),
                         "l"(g.cold), "r"(cb), "r"(bar)
                         : "memory");
        if (lb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base + nb + pb + cb)),
                         "l"(g.lights), "r"(lb), "r"(bar)
                         : "memory");
        if (qb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base + nb + pb + cb + lb)),
                         "l"(g.pairs), "r"(qb), "r"(bar)
                         : "memory");
    }
    uint32_t done = 0; // everyone waits for phase 0 of the barrier
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar)
                     : "memory");
    }
    return acc;
}

// ---- primitive intersection ---------------------------------------------------
// Triangles / parallelograms: the ray is mapped into the primitive's own (b1, b2, h) frame by
// the precomputed affine rows -- the ORIGIN half (three 4-term dot products) is shared by every
// ray that leaves the same point, the DIRECTION half costs 9 FMAs, one fast division and 2 FMAs
// per ray. The accept test is built for the ALU pipe, the busiest one (ncu: 50 % of its peak):
//   * t in (0, tmax): ONE unsigned compare of the float bit patterns -- negative values, -0 and
//     NaN all have larger patterns than any positive bound; a switched-off ray has bound 0
//   * parallelogram (kind 2): the builder centres the coordinates, inside <=> max(|u|,|v|) <= 1/2
//   * triangle (kind 1): inside <=> min(u, v, 1 - u - v) >= 0
// Spheres: unit direction, discriminant from the perpendicular offset; the nearer root ahead of
// the origin is the unsigned minimum of the two roots' bit patterns.
struct PlaneOrigin {
    float ox, oy, oz;
};
__device__ __forceinline__ PlaneOrigin plane_origin(float4 a, float4 b, float4 c, float3 o) {
    PlaneOrigin r;
    r.oz = fmaf(c.x, o.x, fmaf(c.y, o.y, fmaf(c.z, o.z, c.w)));
    r.ox = fmaf(a.x, o.x, fmaf(a.y, o.y, fmaf(a.z, o.z, a.w)));
    r.oy = fmaf(b.x, o.x, fmaf(b.y, o.y, fmaf(b.z, o.z, b.w)));
    return r;
}
// KIND 2 parallelogram, 1 triangle. Returns "hit inside and nearer than lim"; t_bits = pattern of t.
template <int KIND>
__device__ __forceinline__ bool plane_hit(float4 a, float4 b, float4 c, PlaneOrigin po, float3 d, uint32_t lim, uint32_t& t_bits) {
    const float dz = fmaf(c.x, d.x, fmaf(c.y, d.y, c.z * d.z));
    const float t = __fdividef(-po.oz, dz);
    const float dx = fmaf(a.x, d.x, fmaf(a.y, d.y, a.z * d.z));
    const float dy = fmaf(b.x, d.x, fmaf(b.y, d.y, b.z * d.z));
    const float u = fmaf(t, dx, po.ox), v = fmaf(t, dy, po.oy);
    t_bits = __float_as_uint(t);
    if (KIND == 2) return (fmaxf(fabsf(u), fabsf(v)) <= 0.5f) & (t_bits < lim);
    return (fminf(fminf(u, v), 1.0f - (u + v)) >= 0.0f) & (t_bits < lim);
}
__device__ __forceinline__ bool sphere_hit(float3 oc, float radius, float3 d, uint32_t lim, uint32_t& t_bits) {
    const float bq = dot(oc, d);
    const float3 l = oc - d * bq;
    const float disc = fmaf(radius, radius, -dot(l, l));
    const float sq = sqrtf(disc); // NaN when the ray misses: fails the range test below
    t_bits = min(__float_as_uint(-bq - sq), __float_as_uint(sq - bq));
    return t_bits < lim;
}
__device__ __forceinline__ uint32_t range_limit(float tmax) { return tmax > 0.0f ? __float_as_uint(tmax) : 0u; }

// generic form for the mixed leaves of a tree: t in (0, tmax) or -1
__device__ __forceinline__ float hit_prim(float4 a, float4 b, float4 c, float4 tag, float3 o, float3 d, float tmax) {
    uint32_t tb;
    bool ok;
    const uint32_t lim = __float_as_uint(tmax); // callers pass tmax > 0
    if (tag.z != 0.0f) {
        const PlaneOrigin po = plane_origin(a, b, c, o);
        ok = tag.z > 1.5f ? plane_hit<2>(a, b, c, po, d, lim, tb) : plane_hit<1>(a, b, c, po, d, lim, tb);
    } else {
        ok = sphere_hit(o - f3(a.x, a.y, a.z), a.w, d, lim, tb);
    }
    return ok ? __uint_as_float(tb) : -1.0f;
}

// ---- packed FP32 pairs (sm_100: FFMA2 / FMUL2 / FADD2) -----------------------------------
// One issue slot carries two FP32 lanes; an operand built as pk(x, x) compiles to the scalar
// broadcast form (R.F32), so only the primitive coefficients need to be stored as pairs. Measured
// here (tools/ubench/ffma2.cu): 0.98 FFMA or 0.49 FFMA2 per cycle and scheduler, same 4-cycle
// dependent latency -- the FMA pipe does the same work, the ISSUE slots halve, and issue slots are
// what bounds these kernels (ncu: issue 71 %, FMA pipe 34 %, 40 % of the instructions FP32).
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// A ray as the pair loops want it: direction and negated direction as scalars (broadcast operands).
struct RayDir {
    float x, y, z, nx, ny, nz;
};
__device__ __forceinline__ RayDir ray_dir(float3 d) { return RayDir{d.x, d.y, d.z, -d.x, -d.y, -d.z}; }

// Two planar primitives at once. Pair record = 6 x ulonglong2: row a = (x,y | z,w), row b, row c,
// every coefficient a float2 (primitive 2j, primitive 2j+1). Same operation order per lane as
// plane_origin / plane_hit above, so the results are bit-identical to the scalar form.
struct PlaneOrigin2 {
    f2 ox, oy, oz;
};
__device__ __forceinline__ PlaneOrigin2 plane_origin2(const ulonglong2* r, float3 o) {
    const f2 ox = bc(o.x), oy = bc(o.y), oz = bc(o.z);
    const ulonglong2 a0 = r[0], a1 = r[1], b0 = r[2], b1 = r[3], c0 = r[4], c1 = r[5];
    PlaneOrigin2 p;
    p.oz = fma2(c0.x, ox, fma2(c0.y, oy, fma2(c1.x, oz, c1.y)));
    p.ox = fma2(a0.x, ox, fma2(a0.y, oy, fma2(a1.x, oz, a1.y)));
    p.oy = fma2(b0.x, ox, fma2(b0.y, oy, fma2(b1.x, oz, b1.y)));
    return p;
}
// t = -h(o) / h'(d) is formed as h(o) * rcp(h'(-d)): no negation of a pair is needed.
__device__ __forceinline__ void plane_uvt2(const ulonglong2* r, const PlaneOrigin2& po, const RayDir& d, f2& u, f2& v, f2& t) {
    const ulonglong2 a0 = r[0], a1 = r[1], b0 = r[2], b1 = r[3], c0 = r[4], c1 = r[5];
    const f2 ndz = fma2(c0.x, bc(d.nx), fma2(c0.y, bc(d.ny), mul2(c1.x, bc(d.nz))));
    float z0, z1;
    upk(ndz, z0, z1);
    t = mul2(po.oz, pk(rcp_fast(z0), rcp_fast(z1)));
    const f2 dx = fma2(a0.x, bc(d.x), fma2(a0.y, bc(d.y), mul2(a1.x, bc(d.z))));
    const f2 dy = fma2(b0.x, bc(d.x), fma2(b0.y, bc(d.y), mul2(b1.x, bc(d.z))));
    u = fma2(t, dx, po.ox);
    v = fma2(t, dy, po.oy);
}
template <int KIND> __device__ __forceinline__ bool plane_inside(float u, float v) {
    if (KIND == 2) return fmaxf(fabsf(u), fabsf(v)) <= 0.5f;
    return fminf(fminf(u, v), 1.0f - (u + v)) >= 0.0f;
}

// Flat scene (ONE leaf, fully staged, primitives sorted parallelograms | triangles | spheres by
// the builder): nearest hit of ray (o, d1) within tmax1 and, when DUAL, any hit of ray (o, d0)
// within tmax0 -- both leave the same point, so one pass over the primitives serves both. A ray
// with tmax <= 0 is switched off. Three branch-free loops over primitive PAIRS (each kind group is
// padded to an even count with a NaN record that can never be hit); the position in its group IS
// the primitive id.
template <int KIND, bool DUAL>
__device__ __forceinline__ void plane_pairs(const ulonglong2* rec, uint32_t n_pairs, uint32_t id0, float3 o, const RayDir& d1,
                                            const RayDir& d0, uint32_t& lim1, uint32_t& lim0, uint32_t& hit_k) {
#pragma unroll 1
    for (uint32_t j = 0; j < n_pairs; ++j, rec += 6) {
        const PlaneOrigin2 po = plane_origin2(rec, o);
        f2 u, v, t;
        plane_uvt2(rec, po, d1, u, v, t);
        float ua, ub, va, vb, ta, tb;
        upk(u, ua, ub); upk(v, va, vb); upk(t, ta, tb);
        if (plane_inside<KIND>(ua, va) & (__float_as_uint(ta) < lim1)) { lim1 = __float_as_uint(ta); hit_k = id0 + 2u * j; }
        if (plane_inside<KIND>(ub, vb) & (__float_as_uint(tb) < lim1)) { lim1 = __float_as_uint(tb); hit_k = id0 + 2u * j + 1u; }
        if (DUAL) {
            plane_uvt2(rec, po, d0, u, v, t);
            upk(u, ua, ub); upk(v, va, vb); upk(t, ta, tb);
            if ((plane_inside<KIND>(ua, va) & (__float_as_uint(ta) < lim0)) | (plane_inside<KIND>(ub, vb) & (__float_as_uint(tb) < lim0)))
                lim0 = 0u;
        }
    }
}
// Sphere pair record = 2 x ulonglong2: (-cx, -cy | -cz, r^2). Unit direction, discriminant from the
// perpendicular offset; the nearer root ahead of the origin is the unsigned minimum of the roots.
__device__ __forceinline__ void sphere_roots2(f2 ocx, f2 ocy, f2 ocz, f2 rr, const RayDir& d, uint32_t& ta, uint32_t& tb) {
    const f2 bq = fma2(ocz, bc(d.z), fma2(ocy, bc(d.y), mul2(ocx, bc(d.x))));
    const f2 lx = fma2(bq, bc(d.nx), ocx), ly = fma2(bq, bc(d.ny), ocy), lz = fma2(bq, bc(d.nz), ocz);
    const f2 disc = sub2(rr, fma2(lz, lz, fma2(ly, ly, mul2(lx, lx))));
    float b0, b1, q0, q1;
    upk(bq, b0, b1);
    upk(disc, q0, q1);
    const float s0 = sqrtf(q0), s1 = sqrtf(q1); // NaN when the ray misses: fails the range test
    ta = min(__float_as_uint(-b0 - s0), __float_as_uint(s0 - b0));
    tb = min(__float_as_uint(-b1 - s1), __float_as_uint(s1 - b1));
}
template <bool DUAL>
__device__ __forceinline__ void trace_flat(const SceneAccess<true>& S, float3 o, float3 d1, float tmax1, float3 d0,
                                           float tmax0, float& best, uint32_t& best_k, bool& occluded) {
    uint32_t lim1 = range_limit(tmax1), lim0 = DUAL ? range_limit(tmax0) : 0u;
    uint32_t hit_k = kInvalid;
    const uint32_t n_par = uint32_t(S.g->n_par), n_tri = uint32_t(S.g->n_tri), n = uint32_t(S.g->n_prims);
    const uint32_t p_par = (n_par + 1u) >> 1, p_tri = (n_tri + 1u) >> 1, p_sph = (n - n_par - n_tri + 1u) >> 1;
    const RayDir r1 = ray_dir(d1), r0 = ray_dir(d0);
    const ulonglong2* rec = S.pairs_s;
    plane_pairs<2, DUAL>(rec, p_par, 0u, o, r1, r0, lim1, lim0, hit_k);
    rec += 6u * p_par;
    plane_pairs<1, DUAL>(rec, p_tri, n_par, o, r1, r0, lim1, lim0, hit_k);
    rec += 6u * p_tri;
#pragma unroll 1
    for (uint32_t j = 0; j < p_sph; ++j, rec += 2) {
        const ulonglong2 c0 = rec[0], c1 = rec[1];
        const f2 ocx = add2(c0.x, bc(o.x)), ocy = add2(c0.y, bc(o.y)), ocz = add2(c1.x, bc(o.z));
        uint32_t ta, tb;
        sphere_roots2(ocx, ocy, ocz, c1.y, r1, ta, tb);
        const uint32_t id = n_par + n_tri + 2u * j;
        if (ta < lim1) { lim1 = ta; hit_k = id; }
        if (tb < lim1) { lim1 = tb; hit_k = id + 1u; }
        if (DUAL) {
            sphere_roots2(ocx, ocy, ocz, c1.y, r0, ta, tb);
            if ((ta < lim0) | (tb < lim0)) lim0 = 0u;
        }
    }
    best = __uint_as_float(lim1);
    best_k = hit_k;
    occluded = DUAL && tmax0 > 0.0f && lim0 == 0u;
}

// Ordered traversal of the linear octree (tree scenes). Parametric front-to-back walk (after
// Revelles et al. 2000): in the frame where the ray direction is positive on every axis (octant
// bits XOR a), a cell is described per axis by the ray parameters of its entry plane t0, mid
// plane tm and exit plane t1. The first child is the one whose mid planes lie before the entry
// point; the next child is reached through the exit plane with the smallest parameter. Only the
// (at most four) children the ray pierces are visited, strictly in order, so the walk stops at
// the first leaf whose hit lies inside its own cell.
//
// Per level the state is one 4-bit child code in a 64-bit register, the child base index on a
// short stack in shared memory (one column per thread, conflict free) and the integer cell
// coordinates. Plane parameters are linear in the plane index, t(i) = A + i * (B * 2^-(level+1))
// with A = (root_lo - o) / d and B = root_size / d per axis, evaluated with one FMA each: the
// planes a child shares with its parent get bit-identical parameters (same exact product), so
// the walk is consistent from level to level.
//
// What the ncu capture of the 1M-triangle heightfield drove (profiles/r01b_heightfield_*):
//   * 4.4 of 32 threads active per instruction: the one-big-loop form serialised "step",
//     "descend", "test leaf" and "pop" branches -> WHILE-WHILE form: every lane first walks to
//     its next non-empty leaf (lanes that found one wait at the reconvergence point), then all
//     lanes test their leaves together
//   * 443 MB of leaf-ordered primitive copies (6.9 references per primitive) streamed from DRAM
//     -> leaves hold 4-byte indices into the 64-byte records: nodes + indices + records = 108 MB,
//     L2-resident on B200 (126 MB)
//   * 51 instructions per plane update -> 21 with the linear form
// `any` = stop at the first hit (shadow rays).
// The walk as a resumable state machine: init() places the ray at the root, step() advances it
// to its next non-empty leaf and tests that leaf (one while-while round), returning true when
// the ray is finished. traverse() below runs it to completion; trace_kernel interleaves the
// steps of 32 rays per warp and hands a finished lane the next ray of the queue.
struct TreeWalk {
    float3 o, d, A, B;
    float best;
    uint32_t best_prim, a;
    int level;
    uint32_t ix, iy, iz;
    float t0x, t0y, t0z, tmx, tmy, tmz, t1x, t1y, t1z;
    unsigned long long codes; // 4 bits per level: 0xF = not started, else current child (mirrored)
    uint32_t leaf_first, leaf_n; // the part of the current leaf's list still to be tested
    float leaf_exit;             // parameter at which the ray leaves the current leaf's cell
    bool any;

    __device__ __forceinline__ void planes() { // entry / mid / exit parameters of the cell (level; ix,iy,iz)
        const float s1 = __int_as_float((127 - (level + 1)) << 23); // 2^-(level+1), exact
        const float hx = B.x * s1, hy = B.y * s1, hz = B.z * s1;
        const float fx = float(2u * ix), fy = float(2u * iy), fz = float(2u * iz);
        const float ax = fmaf(fx, hx, A.x), bx = fmaf(fx + 2.0f, hx, A.x);
        const float ay = fmaf(fy, hy, A.y), by = fmaf(fy + 2.0f, hy, A.y);
        const float az = fmaf(fz, hz, A.z), bz = fmaf(fz + 2.0f, hz, A.z);
        t0x = fminf(ax, bx); t1x = fmaxf(ax, bx);
        t0y = fminf(ay, by); t1y = fmaxf(ay, by);
        t0z = fminf(az, bz); t1z = fmaxf(az, bz);
        tmx = fmaf(fx + 1.0f, hx, A.x);
        tmy = fmaf(fy + 1.0f, hy, A.y);
        tmz = fmaf(fz + 1.0f, hz, A.z);
    }

    // tests up to `limit` primitives of the list [first, first + n), two per step (independent FMA
    // chains, all loads of the step issued up front)
    __device__ __forceinline__ void leaf(const PathSceneD& g, uint32_t first, uint32_t n, uint32_t limit) {
        const uint32_t* __restrict__ index = g.index;
        const float4* __restrict__ hot = reinterpret_cast<const float4*>(g.hot);
        if (n > limit) n = limit;
        uint32_t id0 = __ldg(index + first), id1 = n > 1 ? __ldg(index + first + 1) : id0;
        for (uint32_t k = 0; k < n; k += 2) {
            const float4* p0 = hot + 4 * (size_t)id0;
            const float4* p1 = hot + 4 * (size_t)id1;
            const float4 a0 = __ldg(p0), b0 = __ldg(p0 + 1), c0 = __ldg(p0 + 2), g0 = __ldg(p0 + 3);
            const float4 a1 = __ldg(p1), b1 = __ldg(p1 + 1), c1 = __ldg(p1 + 2), g1 = __ldg(p1 + 3);
            const uint32_t cur0 = id0, cur1 = id1;
            const bool second = k + 1 < n;
            if (k + 2 < n) id0 = __ldg(index + first + k + 2); // next step's indices while this one computes
            if (k + 3 < n) id1 = __ldg(index + first + k + 3);
            else id1 = id0;
            const float t0 = hit_prim(a0, b0, c0, g0, o, d, best);
            if (t0 >= 0.0f) { best = t0; best_prim = cur0; }
            const float t1 = second ? hit_prim(a1, b1, c1, g1, o, d, best) : -1.0f;
            if (t1 >= 0.0f) { best = t1; best_prim = cur1; }
            if (any && best_prim != kInvalid) break;
        }
    }

    // returns true when the ray is already finished (missed the root box, or the scene is one leaf)
    template <bool ALL>
    __device__ __forceinline__ bool init(const SceneAccess<ALL>& S, float3 o_, float3 d_, float tmax, bool any_) {
        const PathSceneD& g = *S.g;
        o = o_;
        d = d_;
        best = tmax;
        best_prim = kInvalid;
        any = any_;
        const uint2 root = S.node(0);
        leaf_n = 0;
        if (root.y & kLeafBit) { // the whole scene is one leaf (but not staged as a flat scene)
            leaf(g, root.x, root.y & ~kLeafBit, 0x7fffffffu);
            return true;
        }
        float3 dd = d;
        if (fabsf(dd.x) < 1.0e-20f) dd.x = copysignf(1.0e-20f, dd.x);
        if (fabsf(dd.y) < 1.0e-20f) dd.y = copysignf(1.0e-20f, dd.y);
        if (fabsf(dd.z) < 1.0e-20f) dd.z = copysignf(1.0e-20f, dd.z);
        const float3 inv = f3(__fdividef(1.0f, dd.x), __fdividef(1.0f, dd.y), __fdividef(1.0f, dd.z));
        a = (dd.x < 0.0f ? 1u : 0u) | (dd.y < 0.0f ? 2u : 0u) | (dd.z < 0.0f ? 4u : 0u);
        A = f3((g.root_lo[0] - o.x) * inv.x, (g.root_lo[1] - o.y) * inv.y, (g.root_lo[2] - o.z) * inv.z);
        B = f3(g.root_size[0] * inv.x, g.root_size[1] * inv.y, g.root_size[2] * inv.z);
        level = 0;
        ix = iy = iz = 0;
        planes();
        const float tn = fmaxf(fmaxf(t0x, t0y), t0z), tf = fminf(fminf(t1x, t1y), t1z);
        if (tn > fminf(tf, best) + 1.0e-5f || tf < 0.0f) return true;
        codes = 0xFull;
        S.stack[0] = root.x;
        return false;
    }

    // One ROUND, built so that the 32 rays of a warp stay in step: exactly kWalkSteps cell moves
    // towards the next non-empty leaf (predicated off for lanes that have arrived or are finished),
    // then one batch of at most kLeafBatch primitives of the current leaf. MUST be called by all 32
    // lanes of the warp (`active` = this lane has a ray): the loop re-converges the warp at every
    // move -- with `continue`/`break` in a data-dependent loop the lanes drifted apart and ran the
    // loop body in ~2.5 separate groups of 5.6 lanes (ncu, profiles/r01_tuning_log.md).
    // Returns true when the lane's ray is finished.
    template <bool ALL, bool COOP> __device__ __forceinline__ bool step(const SceneAccess<ALL>& S, bool active) {
        uint32_t* const stack = S.stack;
        bool done = false;
        const int kWalkSteps = S.walk_steps;      // PassArgs::walk_steps (default 4)
        const uint32_t kLeafBatch = S.leaf_batch; // PassArgs::leaf_batch (default 4)
#pragma unroll 1
        for (int it = 0; it < kWalkSteps; ++it) {
            __syncwarp();
            if (active && !done && leaf_n == 0u) {
                // first child / next child, both evaluated branch-free (a branch here made the
                // compiler run the rest of the move once per side: ncu, 2 x 5.6 lanes)
                const uint32_t old = uint32_t(codes >> (4 * level)) & 0xFu;
                const bool fresh = old == 0xFu;
                const float te = fmaxf(fmaxf(fmaxf(t0x, t0y), t0z), 0.0f);
                const uint32_t first = (tmx < te ? 1u : 0u) | (tmy < te ? 2u : 0u) | (tmz < te ? 4u : 0u);
                const float ex = (old & 1u) ? t1x : tmx, ey = (old & 2u) ? t1y : tmy, ez = (old & 4u) ? t1z : tmz;
                const uint32_t bit = (ex <= ey && ex <= ez) ? 1u : (ey <= ez ? 2u : 4u);
                const bool leave = !fresh && (old & bit) != 0u;
                const uint32_t cur = fresh ? first : (old | bit);
                bool moved = false;
                if (leave) { // the ray left this cell: back to the parent
                    if (level == 0) {
                        done = true;
                    } else {
                        --level;
                        ix >>= 1; iy >>= 1; iz >>= 1;
                        moved = true;
                    }
                } else {
                    const float cen = fmaxf(fmaxf((cur & 1u) ? tmx : t0x, (cur & 2u) ? tmy : t0y), (cur & 4u) ? tmz : t0z);
                    const float cex = fminf(fminf((cur & 1u) ? t1x : tmx, (cur & 2u) ? t1y : tmy), (cur & 4u) ? t1z : tmz);
                    if (cen > best + fabsf(best) * 2.0e-6f + 1.0e-5f) {
                        done = true; // everything from here on is farther than the hit
                    } else {
                        codes = (codes & ~(0xFull << (4 * level))) | ((unsigned long long)cur << (4 * level));
                        if (cex >= 0.0f) {
                            const uint32_t c = cur ^ a;
                            const uint2 rec = S.node(stack[level * kThreads] + c);
                            if (rec.y != kLeafBit) { // not an empty octant
                                if (rec.y & kLeafBit) {
                                    leaf_first = rec.x;
                                    leaf_n = rec.y & ~kLeafBit;
                                    leaf_exit = cex;
                                } else {
                                    ++level;
                                    stack[level * kThreads] = rec.x;
                                    ix = 2u * ix + (c & 1u); iy = 2u * iy + ((c >> 1) & 1u); iz = 2u * iz + ((c >> 2) & 1u);
                                    codes |= 0xFull << (4 * level);
                                    moved = true;
                                }
                            }
                        }
                    }
                }
                if (moved) planes();
            }
        }
        __syncwarp();
        // ---- test: the next batch of the current leaf's primitives ----
        if constexpr (COOP) {
            // COOPERATIVE form (default; G19_COOP_LEAF=0 selects the sequential one): ncu showed the leaf batches running at ~8 of 32 lanes (only
            // the lanes that stand in a leaf) for 28 % of trace_kernel's instructions. Here the warp's pending
            // (ray, primitive) pairs are spread over all 32 lanes: a lane looks up which ray task j belongs to
            // (binary search over the inclusive prefix sum of the batch sizes), fetches that ray by shuffle,
            // tests ONE primitive, and the nearest hit per ray is reduced with a 64-bit atomicMin on
            // (t bits << 32 | primitive) in the warp's shared-memory slots. Same hits as the sequential form;
            // a tie in t goes to the lower primitive id.
            const uint32_t lane = threadIdx.x & 31u;
            const bool mine = active && !done && leaf_n != 0u;
            const uint32_t cnt = mine ? (leaf_n < kLeafBatch ? leaf_n : kLeafBatch) : 0u;
            if (__ballot_sync(kFull, cnt != 0u) != 0u) {
                uint32_t inc = cnt; // inclusive prefix sum over the lanes
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t v = __shfl_up_sync(kFull, inc, off);
                    if (lane >= uint32_t(off)) inc += v;
                }
                const uint32_t total = __shfl_sync(kFull, inc, 31);
                const uint32_t exc = inc - cnt;
                S.coop[lane] = ~0ull;
                __syncwarp();
                const uint32_t* __restrict__ index = S.g->index;
                const float4* __restrict__ hot = reinterpret_cast<const float4*>(S.g->hot);
                for (uint32_t base = 0; base < total; base += 32u) {
                    const uint32_t j = base + lane;
                    uint32_t own = 0; // number of lanes whose tasks all lie before j = the owner of task j
#pragma unroll
                    for (int stepw = 16; stepw > 0; stepw >>= 1) {
                        const uint32_t probe = __shfl_sync(kFull, inc, int(own) + stepw - 1);
                        if (probe <= j) own += uint32_t(stepw);
                    }
                    const bool valid = j < total;
                    if (!valid) own = 0;
                    const uint32_t k = j - __shfl_sync(kFull, exc, int(own));
                    const uint32_t lf = __shfl_sync(kFull, leaf_first, int(own));
                    const float rox = __shfl_sync(kFull, o.x, int(own)), roy = __shfl_sync(kFull, o.y, int(own)), roz = __shfl_sync(kFull, o.z, int(own));
                    const float rdx = __shfl_sync(kFull, d.x, int(own)), rdy = __shfl_sync(kFull, d.y, int(own)), rdz = __shfl_sync(kFull, d.z, int(own));
                    const float rbest = __shfl_sync(kFull, best, int(own));
                    if (valid) {
                        const uint32_t id = __ldg(index + lf + k);
                        const float4* pp = hot + 4 * (size_t)id;
                        const float4 a0 = __ldg(pp), b0 = __ldg(pp + 1), c0 = __ldg(pp + 2), g0 = __ldg(pp + 3);
                        const float t = hit_prim(a0, b0, c0, g0, f3(rox, roy, roz), f3(rdx, rdy, rdz), rbest);
                        if (t >= 0.0f) atomicMin(S.coop + own, ((unsigned long long)__float_as_uint(t) << 32) | id);
                    }
                }
                __syncwarp();
                if (cnt != 0u) {
                    const unsigned long long r = S.coop[lane];
                    if (r != ~0ull) { best = __uint_as_float(uint32_t(r >> 32)); best_prim = uint32_t(r); }
                    leaf_first += cnt;
                    leaf_n -= cnt;
                    if (any && best_prim != kInvalid) done = true;
                    else if (leaf_n == 0u && best_prim != kInvalid && best <= leaf_exit) done = true;
                }
                __syncwarp();
            }
        } else if (active && !done && leaf_n != 0u) { // sequential form: every lane tests its own leaf's primitives
            leaf(*S.g, leaf_first, leaf_n, kLeafBatch);
            const uint32_t tested = leaf_n < kLeafBatch ? leaf_n : kLeafBatch;
            leaf_first += tested;
            leaf_n -= tested;
            if (any && best_prim != kInvalid) done = true;                                // any hit ends a shadow ray
            else if (leaf_n == 0u && best_prim != kInvalid && best <= leaf_exit) done = true; // a hit inside its own cell cannot be beaten
        }
        __syncwarp();
        return done;
    }
};

// ---- TreeWalk2: point-location restart walk ---------------------------------------------------
// What the round-1 capture of the parametric walk above showed (profiles/r01f_tree_kernels_ncu.json): ~870 warp
// instructions per ray at 13.7 of 32 lanes, 62 % of them cell MOVES at ~130 instructions each -- every move
// recomputes nine plane parameters, edits a 64-bit code word with variable shifts and runs a four-way
// leave / descend / skip / arrive decision, and going from one leaf to a neighbour under another ancestor costs one
// move per level up plus one per level down. This walk keeps the ray's position as INTEGER coordinates (X, Y, Z) on
// the finest grid of the octree (2^14 cells per axis; node boxes are implicit, so a cell of level l is the top l
// bits) and does three cheap things per visited cell:
//   descend   from the deepest ancestor still valid for (X, Y, Z) -- its child base sits on the per-thread stack in
//             shared memory -- follow the coordinate bits down to the leaf or empty cell: three bit extracts, one
//             8-byte node record, ~12 instructions per level, typically 1-3 levels
//   exit      the parameter at which the ray leaves the cell: one FMA per axis on the same linear plane form
//             t(i) = A + i * (B * 2^-level) as before (shared planes get bit-identical parameters), minimum of three
//   advance   step the exit axis' coordinate to the neighbouring cell (exact integer arithmetic: no epsilon nudging),
//             re-derive the other two from the exit point clamped to the current cell's range, and find the common
//             ancestor with one XOR + CLZ
// -- no plane registers, no code word, no pops. Cells are visited strictly in ray order, so the walk still stops at
// the first leaf whose hit lies inside its own cell. COUNT: node records visited and primitives tested are tallied
// for the issue-roofline model; instantiated only under params.profile.
template <bool COUNT> struct TreeWalk2 {
    float3 o, d, A, B;
    float best;
    uint32_t best_prim;
    uint32_t X, Y, Z;            // position on the finest grid
    uint32_t lv;                 // deepest level whose stack entry (child base) is valid for (X, Y, Z)
    float t_exit;                // where the ray leaves the current cell (valid while an advance is pending)
    uint32_t leaf_first, leaf_n; // the part of the current leaf's list still to be tested
    float leaf_exit;
    uint32_t state;              // bit 0: any-hit ray; bit 1: advance pending; bits 2-3: exit axis
    uint32_t n_node, n_prim;     // COUNT only

    template <bool ALL>
    __device__ __forceinline__ bool init(const SceneAccess<ALL>& S, float3 o_, float3 d_, float tmax, bool any_) {
        const PathSceneD& g = *S.g;
        o = o_;
        d = d_;
        best = tmax;
        best_prim = kInvalid;
        state = any_ ? 1u : 0u;
        leaf_n = 0;
        const uint2 root = S.node(0);
        if (root.y & kLeafBit) { // the whole scene is one leaf (but not staged as a flat scene)
            TreeWalk w0 = {};
            w0.o = o; w0.d = d; w0.best = best; w0.best_prim = kInvalid; w0.any = any_;
            w0.leaf(g, root.x, root.y & ~kLeafBit, 0x7fffffffu);
            best = w0.best;
            best_prim = w0.best_prim;
            return true;
        }
        // slab clip against the root box; an axis the ray (almost) does not move along never produces an exit
        float tn = 0.0f, tf = best;
        bool miss = false;
        const float inf = __int_as_float(0x7f800000);
#define G19_AXIS(c, k)                                                                  \
        if (fabsf(d.c) < 1.0e-12f) {                                                    \
            miss |= o.c < g.root_lo[k] || o.c > g.root_lo[k] + g.root_size[k];          \
            A.c = inf; B.c = 0.0f;                                                      \
        } else {                                                                        \
            const float inv = __fdividef(1.0f, d.c);                                    \
            A.c = (g.root_lo[k] - o.c) * inv;                                           \
            B.c = g.root_size[k] * inv;                                                 \
            const float t1 = A.c + B.c;                                                 \
            tn = fmaxf(tn, fminf(A.c, t1));                                             \
            tf = fminf(tf, fmaxf(A.c, t1));                                             \
        }
        G19_AXIS(x, 0) G19_AXIS(y, 1) G19_AXIS(z, 2)
#undef G19_AXIS
        if (miss || tn > tf + 1.0e-5f) return true;
        const int hi = (1 << kMaxTreeDepth) - 1;
        X = uint32_t(min(max(__float2int_rd((fmaf(d.x, tn, o.x) - g.root_lo[0]) * g.grid_scale[0]), 0), hi));
        Y = uint32_t(min(max(__float2int_rd((fmaf(d.y, tn, o.y) - g.root_lo[1]) * g.grid_scale[1]), 0), hi));
        Z = uint32_t(min(max(__float2int_rd((fmaf(d.z, tn, o.z) - g.root_lo[2]) * g.grid_scale[2]), 0), hi));
        lv = 0;
        S.stack[0] = root.x;
        if (COUNT) n_node = n_prim = 0;
        return false;
    }

    // One ROUND for the 32 rays of a warp (MUST be called by all lanes, `active` = this lane has a ray): walk_steps
    // times [advance out of the finished cell, descend to the next leaf / empty cell, compute its exit], then one
    // batch of at most leaf_batch primitives of the lanes that stand in a leaf. Returns true when the lane's ray is
    // finished.
    template <bool ALL, bool COOP> __device__ __forceinline__ bool step(const SceneAccess<ALL>& S, bool active) {
        uint32_t* const stack = S.stack;
        const PathSceneD& g = *S.g;
        bool done = false;
        const int kWalkSteps = S.walk_steps;
        const uint32_t kLeafBatch = S.leaf_batch;
#pragma unroll 1
        for (int it = 0; it < kWalkSteps; ++it) {
            __syncwarp();
            bool walking = active && !done && leaf_n == 0u;
            if (walking && (state & 2u)) {
                // ---- advance: out of the current cell through its exit face ----
                const uint32_t s = uint32_t(kMaxTreeDepth - 1) - lv; // the cell is at level lv + 1: it spans 2^s fine cells per axis
                const uint32_t axis = (state >> 2) & 3u;
                const uint32_t mask = (1u << s) - 1u;
                const int hi = (1 << kMaxTreeDepth) - 1;
                // the exit point on the finest grid, held inside the current cell (it IS on the cell's boundary)
                int nx = __float2int_rd((fmaf(d.x, t_exit, o.x) - g.root_lo[0]) * g.grid_scale[0]);
                int ny = __float2int_rd((fmaf(d.y, t_exit, o.y) - g.root_lo[1]) * g.grid_scale[1]);
                int nz = __float2int_rd((fmaf(d.z, t_exit, o.z) - g.root_lo[2]) * g.grid_scale[2]);
                const int x0 = int(X & ~mask), y0 = int(Y & ~mask), z0 = int(Z & ~mask);
                nx = min(max(nx, x0), x0 + int(mask));
                ny = min(max(ny, y0), y0 + int(mask));
                nz = min(max(nz, z0), z0 + int(mask));
                // ... and one cell further along the exit axis
                if (axis == 0u) nx = d.x > 0.0f ? x0 + int(mask) + 1 : x0 - 1;
                else if (axis == 1u) ny = d.y > 0.0f ? y0 + int(mask) + 1 : y0 - 1;
                else nz = d.z > 0.0f ? z0 + int(mask) + 1 : z0 - 1;
                if ((nx | ny | nz) < 0 || nx > hi || ny > hi || nz > hi) {
                    done = true; // left the root box
                    walking = false;
                } else {
                    const uint32_t diff = (X ^ uint32_t(nx)) | (Y ^ uint32_t(ny)) | (Z ^ uint32_t(nz));
                    // highest differing bit h = 31 - clz: the level-m ancestors agree for m <= kMaxTreeDepth - 1 - h
                    lv = min(lv, uint32_t(__clz(int(diff))) - uint32_t(32 - kMaxTreeDepth));
                    X = uint32_t(nx); Y = uint32_t(ny); Z = uint32_t(nz);
                }
                state &= ~2u;
            }
            // ---- descend to the leaf / empty cell that holds (X, Y, Z) ----
            bool desc = walking;
            uint32_t k = lv;
            uint2 rec = make_uint2(0u, kLeafBit);
            // a descent that would start above the top table's level starts IN the table: one load instead of a
            // chain of top_level dependent ones
            const uint32_t top_level = uint32_t(g.top_level);
            if (desc && k < top_level) {
                const uint32_t sh = uint32_t(kMaxTreeDepth) - top_level;
                rec = __ldg(g.top + ((((size_t(Z >> sh) << top_level) | (Y >> sh)) << top_level) | (X >> sh)));
                if (COUNT) ++n_node;
                if (rec.y & kLeafBit) {
                    k = ((rec.y >> 24) & 0xfu) - 1u; // the covering leaf / empty cell sits at that (coarser) level
                    rec.y &= 0x80ffffffu;
                    desc = false;
                } else {
                    k = top_level;
                    stack[k * kThreads] = rec.x;
                }
            }
            while (__any_sync(kFull, desc)) {
                if (desc) {
                    const uint32_t sh = uint32_t(kMaxTreeDepth - 1) - k;
                    const uint32_t c = ((X >> sh) & 1u) | (((Y >> sh) & 1u) << 1) | (((Z >> sh) & 1u) << 2);
                    rec = S.node(stack[k * kThreads] + c);
                    if (COUNT) ++n_node;
                    if ((rec.y & kLeafBit) || k + 1u >= uint32_t(kMaxTreeDepth)) {
                        desc = false;
                    } else {
                        ++k;
                        stack[k * kThreads] = rec.x;
                    }
                }
            }
            if (walking) {
                lv = k;
                // ---- exit parameter of the cell (level lv + 1): the far plane per axis, linear in the plane index ----
                const uint32_t sh = uint32_t(kMaxTreeDepth - 1) - k;
                const float s1 = __int_as_float((127 - int(k + 1u)) << 23); // 2^-(level), exact
                const float fx = float((X >> sh) + (d.x > 0.0f ? 1u : 0u)), fy = float((Y >> sh) + (d.y > 0.0f ? 1u : 0u)),
                            fz = float((Z >> sh) + (d.z > 0.0f ? 1u : 0u));
                const float tx = fmaf(fx, B.x * s1, A.x), ty = fmaf(fy, B.y * s1, A.y), tz = fmaf(fz, B.z * s1, A.z);
                const uint32_t axis = (tx <= ty && tx <= tz) ? 0u : (ty <= tz ? 1u : 2u);
                t_exit = fminf(fminf(tx, ty), tz);
                state = (state & 1u) | 2u | (axis << 2);
                if (rec.y != kLeafBit) { // a leaf with primitives
                    leaf_first = rec.x;
                    leaf_n = rec.y & ~kLeafBit;
                    leaf_exit = t_exit;
                } else if (t_exit >= best) {
                    done = true; // empty cell, and everything beyond it is farther than the hit / the ray's end
                }
            }
        }
        __syncwarp();
        // ---- test: the next batch of the current leaf's primitives (same two forms as TreeWalk::step) ----
        const bool any = (state & 1u) != 0u;
        if constexpr (COOP) {
            const uint32_t lane = threadIdx.x & 31u;
            const bool mine = active && !done && leaf_n != 0u;
            const uint32_t cnt = mine ? (leaf_n < kLeafBatch ? leaf_n : kLeafBatch) : 0u;
            if (__ballot_sync(kFull, cnt != 0u) != 0u) {
                uint32_t inc = cnt; // inclusive prefix sum over the lanes
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t v = __shfl_up_sync(kFull, inc, off);
                    if (lane >= uint32_t(off)) inc += v;
                }
                const uint32_t total = __shfl_sync(kFull, inc, 31);
                const uint32_t exc = inc - cnt;
                S.coop[lane] = ~0ull;
                __syncwarp();
                const uint32_t* __restrict__ index = g.index;
                const float4* __restrict__ hot = reinterpret_cast<const float4*>(g.hot);
                for (uint32_t base = 0; base < total; base += 32u) {
                    const uint32_t j = base + lane;
                    uint32_t own = 0; // number of lanes whose tasks all lie before j = the owner of task j
#pragma unroll
                    for (int stepw = 16; stepw > 0; stepw >>= 1) {
                        const uint32_t probe = __shfl_sync(kFull, inc, int(own) + stepw - 1);
                        if (probe <= j) own += uint32_t(stepw);
                    }
                    const bool valid = j < total;
                    if (!valid) own = 0;
                    const uint32_t kk = j - __shfl_sync(kFull, exc, int(own));
                    const uint32_t lf = __shfl_sync(kFull, leaf_first, int(own));
                    const float rox = __shfl_sync(kFull, o.x, int(own)), roy = __shfl_sync(kFull, o.y, int(own)), roz = __shfl_sync(kFull, o.z, int(own));
                    const float rdx = __shfl_sync(kFull, d.x, int(own)), rdy = __shfl_sync(kFull, d.y, int(own)), rdz = __shfl_sync(kFull, d.z, int(own));
                    const float rbest = __shfl_sync(kFull, best, int(own));
                    if (valid) {
                        const uint32_t id = __ldg(index + lf + kk);
                        const float4* pp = hot + 4 * (size_t)id;
                        const float4 a0 = __ldg(pp), b0 = __ldg(pp + 1), c0 = __ldg(pp + 2), g0 = __ldg(pp + 3);
                        const float t = hit_prim(a0, b0, c0, g0, f3(rox, roy, roz), f3(rdx, rdy, rdz), rbest);
                        if (t >= 0.0f) atomicMin(S.coop + own, ((unsigned long long)__float_as_uint(t) << 32) | id);
                    }
                }
                __syncwarp();
                if (cnt != 0u) {
                    const unsigned long long r = S.coop[lane];
                    if (r != ~0ull) { best = __uint_as_float(uint32_t(r >> 32)); best_prim = uint32_t(r); }
                    if (COUNT) n_prim += cnt;
                    leaf_first += cnt;
                    leaf_n -= cnt;
                    if (any && best_prim != kInvalid) done = true;
                    else if (leaf_n == 0u && leaf_exit >= best) done = true; // the hit lies inside this cell / the ray ends here
                }
                __syncwarp();
            }
        } else if (active && !done && leaf_n != 0u) { // sequential form: every lane tests its own leaf's primitives
            TreeWalk w0 = {};
            w0.o = o; w0.d = d; w0.best = best; w0.best_prim = best_prim; w0.any = any;
            w0.leaf(g, leaf_first, leaf_n, kLeafBatch);
            best = w0.best;
            best_prim = w0.best_prim;
            const uint32_t tested = leaf_n < kLeafBatch ? leaf_n : kLeafBatch;
            if (COUNT) n_prim += tested;
            leaf_first += tested;
            leaf_n -= tested;
            if (any && best_prim != kInvalid) done = true;
            else if (leaf_n == 0u && leaf_exit >= best) done = true;
        }
        __syncwarp();
        return done;
    }
};

// ---- BvhWalk: bounding-volume hierarchy (bvh_build.cu) --------------------------------------------------
// Same resumable interface as the octree walks (init / step / best / best_prim), so trace_kernel's dynamic ray fetch and
// the camera-ray kernel use it unchanged. A round = walk_steps node visits (a node record is both children's boxes: two
// slab tests against [0, nearest hit so far], descend into the nearer child that is hit, push the other) followed by one
// leaf visit for the lanes that stand at a leaf (1-4 primitives, contiguous 64-byte records in leaf order: one load
// chain). Fixed-trip and predicated like the octree rounds, so the warp re-converges after every visit. The stack of
// postponed children lives in local memory (L1): an LBVH over 30-bit Morton codes is at most ~40 levels deep.
// The few primitives the host kept out of the hierarchy (walls around a mesh: their boxes would cover every level above
// them) are tested first -- which also gives the walk a finite nearest-hit bound from its first node on.
// The primitives the host kept out of the hierarchy, nearest hit within `best`: the flat scenes' pair loops (two
// primitives per packed-FP32 instruction, kind-sorted, no branch per primitive) over records every lane reads from the
// same addresses. Round 2 capture: the room's seven big primitives were 2/3 of all primitive tests of the BVH walk.
__device__ __forceinline__ void bvh_big_hits(const PathSceneD& g, float3 o, float3 d, float& best, uint32_t& best_prim) {
    if (g.n_big == 0) return;
    uint32_t lim1 = range_limit(best), lim0 = 0u, hit_k = kInvalid;
    const uint32_t n_par = uint32_t(g.big_par), n_tri = uint32_t(g.big_tri), n = uint32_t(g.n_big);
    const uint32_t p_par = (n_par + 1u) >> 1, p_tri = (n_tri + 1u) >> 1, p_sph = (n - n_par - n_tri + 1u) >> 1;
    const RayDir r1 = ray_dir(d);
    const ulonglong2* rec = reinterpret_cast<const ulonglong2*>(g.bvh_big_pairs);
    plane_pairs<2, false>(rec, p_par, 0u, o, r1, r1, lim1, lim0, hit_k);
    rec += 6u * p_par;
    plane_pairs<1, false>(rec, p_tri, n_par, o, r1, r1, lim1, lim0, hit_k);
    rec += 6u * p_tri;
#pragma unroll 1
    for (uint32_t j = 0; j < p_sph; ++j, rec += 2) {
        const ulonglong2 c0 = rec[0], c1 = rec[1];
        const f2 ocx = add2(c0.x, bc(o.x)), ocy = add2(c0.y, bc(o.y)), ocz = add2(c1.x, bc(o.z));
        uint32_t ta, tb;
        sphere_roots2(ocx, ocy, ocz, c1.y, r1, ta, tb);
        const uint32_t id = n_par + n_tri + 2u * j;
        if (ta < lim1) { lim1 = ta; hit_k = id; }
        if (tb < lim1) { lim1 = tb; hit_k = id + 1u; }
    }
    if (hit_k != kInvalid) {
        best = __uint_as_float(lim1);
        best_prim = __ldg(g.bvh_big + hit_k);
    }
}

constexpr uint32_t kBvhLeaf = 0x80000000u, kBvhDone = 0xffffffffu;
// Entries beyond the shared-memory levels live in local memory (rarely touched). Together they can hold the deepest walk
// there is: Karras' tree over (30-bit code, 32-bit position) keys is at most 62 levels deep, the binary walk postpones at
// most one child per level, the 4-wide walk (31 levels) at most three.
constexpr int kBvhLocalStack = 96;
struct BvhWalk {
    float3 o, d, idir, ood;
    float best;
    uint32_t best_prim;
    uint32_t cur;  // node index, kBvhLeaf | (count - 1) << 28 | first, or kBvhDone
    uint32_t pend; // a leaf reached but not tested yet (speculative traversal), 0 = none
    uint32_t any;
    int sp, cap; // cap: stack levels in shared memory
    uint32_t stk[kBvhLocalStack];

    // The postponed children: the first kBvhSmemStack levels in the thread's column of shared memory (the capture of the
    // all-local form had 22 % of trace_kernel's stall samples on the push / pop lines), deeper ones in local memory.
    __device__ __forceinline__ void push(uint32_t* __restrict__ smem, uint32_t v) {
        if (sp < cap) smem[sp * kThreads] = v;
        else if (sp < cap + kBvhLocalStack) stk[sp - cap] = v;
        else return;
        ++sp;
    }
    __device__ __forceinline__ void pop(const uint32_t* __restrict__ smem) {
        if (sp == 0) { cur = kBvhDone; return; }
        --sp;
        cur = sp < cap ? smem[sp * kThreads] : stk[sp - cap];
    }

    template <bool ALL>
    __device__ __forceinline__ bool init(const SceneAccess<ALL>& S, float3 o_, float3 d_, float tmax, bool any_) {
        const PathSceneD& g = *S.g;
        o = o_;
        d = d_;
        best = tmax;
        best_prim = kInvalid;
        any = any_ ? 1u : 0u;
        sp = 0;
        cap = S.bvh_stack;
        pend = 0u;
        bvh_big_hits(g, o, d, best, best_prim);
        if (any_ && best_prim != kInvalid) return true;
        float3 dd = d;
        if (fabsf(dd.x) < 1.0e-20f) dd.x = copysignf(1.0e-20f, dd.x);
        if (fabsf(dd.y) < 1.0e-20f) dd.y = copysignf(1.0e-20f, dd.y);
        if (fabsf(dd.z) < 1.0e-20f) dd.z = copysignf(1.0e-20f, dd.z);
        idir = f3(__fdividef(1.0f, dd.x), __fdividef(1.0f, dd.y), __fdividef(1.0f, dd.z));
        ood = f3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        cur = g.bvh_root;
        return cur == kBvhDone;
    }

    template <bool ALL, bool COOP> __device__ __forceinline__ bool step(const SceneAccess<ALL>& S, bool active) {
        const PathSceneD& g = *S.g;
        const int kWalkSteps = S.walk_steps;
        const float4* __restrict__ nodes = g.bvh_nodes;
#pragma unroll 1
        const bool spec = (S.bvh_spec & 1) != 0;
        for (int it = 0; it < kWalkSteps; ++it) {
            __syncwarp();
            // Speculative traversal (Aila & Laine 2009): a lane that arrives at a leaf while the others still descend
            // puts the leaf aside and goes on with its next postponed node instead of idling until the leaf step. The
            // nearest-hit bound is then one leaf behind, which can only cost visits, never a hit.
            if (spec && active && pend == 0u && cur >= kBvhLeaf && cur != kBvhDone) {
                pend = cur;
                pop(S.stack);
            }
            if (active && cur < kBvhLeaf) { // an internal node: both children's slabs
                const float4* __restrict__ np = nodes + 4 * (size_t)cur;
                const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
                // slightly widened: a hit ON a box face must not be lost to rounding (the primitive test decides)
                const float hi = best * 1.000002f + 1.0e-6f;
                const float c0lox = fmaf(n0.x, idir.x, -ood.x), c0hix = fmaf(n0.y, idir.x, -ood.x);
                const float c0loy = fmaf(n0.z, idir.y, -ood.y), c0hiy = fmaf(n0.w, idir.y, -ood.y);
                const float c0loz = fmaf(n2.x, idir.z, -ood.z), c0hiz = fmaf(n2.y, idir.z, -ood.z);
                const float c1lox = fmaf(n1.x, idir.x, -ood.x), c1hix = fmaf(n1.y, idir.x, -ood.x);
                const float c1loy = fmaf(n1.z, idir.y, -ood.y), c1hiy = fmaf(n1.w, idir.y, -ood.y);
                const float c1loz = fmaf(n2.z, idir.z, -ood.z), c1hiz = fmaf(n2.w, idir.z, -ood.z);
                const float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), 0.0f));
                const float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), hi));
                const float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), 0.0f));
                const float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), hi));
                // (1 + 3 ulp on the far side: the products above round independently)
                const bool h0 = t0n <= t0f * 1.0000004f, h1 = t1n <= t1f * 1.0000004f;
                const uint32_t r0 = __float_as_uint(n3.x), r1 = __float_as_uint(n3.y);
                const bool both = h0 && h1, swap = both && t1n < t0n; // swap: child 1 is entered first
                if (both) push(S.stack, swap ? r0 : r1);
                if (h0 || h1) cur = (h0 && !swap) ? r0 : r1;
                else pop(S.stack);
            }
        }
        __syncwarp();
        if (active && pend == 0u && cur >= kBvhLeaf && cur != kBvhDone) { // the leaf the lane stands at
            pend = cur;
            pop(S.stack);
        }
        if (active && pend != 0u) { // a leaf: its primitives, nearest so far kept
            const uint32_t first = pend & 0x0fffffffu, cnt = ((pend >> 28) & 7u) + 1u;
            pend = 0u;
            const float4* __restrict__ pp = g.bvh_prims + 4 * (size_t)first;
#pragma unroll 1
            for (uint32_t k = 0; k < cnt; ++k, pp += 4) {
                const float4 a0 = __ldg(pp), b0 = __ldg(pp + 1), c0 = __ldg(pp + 2), g0 = __ldg(pp + 3);
                const float t = hit_prim(a0, b0, c0, g0, o, d, best);
                if (t >= 0.0f) { best = t; best_prim = __float_as_uint(g0.w); }
            }
            if (any && best_prim != kInvalid) cur = kBvhDone;
        }
        __syncwarp();
        return active && cur == kBvhDone;
    }
};

// ---- Bvh4Walk: the 4-wide, quantised form of the same hierarchy ---------------------------------------------
// What the capture of the binary walk showed (profiles/r02m_*): issue 37 %, but the L1 data pipe at 81 % of its peak --
// every lane pulls its own 64-byte node record through L1, four wavefronts per visit. Here the same four wavefronts
// carry FOUR children (boxes as 16-bit coordinates on the root box's grid, bvh_build.cu), and a ray makes half as many
// dependent visits. A coordinate q becomes a float with one byte permute and one subtraction (0x4B000000 | q is the
// float 2^23 + q), the grid scale and the ray's origin are folded into the per-ray slab constants (sc, of), the hit
// children are ordered by entry distance with a five-exchange network on (distance bits | slot) keys, nearest entered,
// the others pushed farthest first.
template <bool COUNT> struct Bvh4Walk { // COUNT: box and primitive tests are tallied (params.profile: the issue-roofline model)
    float3 o, d, sc, of; // slab parameter of grid coordinate q along x: q * sc.x + of.x
    uint32_t n_node, n_prim; // COUNT only
    float best;
    uint32_t best_prim;
    uint32_t cur, pend, any;
    int sp, cap; // cap: stack levels in shared memory
    uint32_t stk[kBvhLocalStack];

    __device__ __forceinline__ void push(uint32_t* __restrict__ smem, uint32_t v) {
        if (sp < cap) smem[sp * kThreads] = v;
        else if (sp < cap + kBvhLocalStack) stk[sp - cap] = v;
        else return;
        ++sp;
    }
    __device__ __forceinline__ void pop(const uint32_t* __restrict__ smem) {
        if (sp == 0) { cur = kBvhDone; return; }
        --sp;
        cur = sp < cap ? smem[sp * kThreads] : stk[sp - cap];
    }

    template <bool ALL>
    __device__ __forceinline__ bool init(const SceneAccess<ALL>& S, float3 o_, float3 d_, float tmax, bool any_) {
        const PathSceneD& g = *S.g;
        o = o_;
        d = d_;
        best = tmax;
        best_prim = kInvalid;
        any = any_ ? 1u : 0u;
        sp = 0;
        cap = S.bvh_stack;
        pend = 0u;
        if (COUNT) { n_node = 0; n_prim = uint32_t(g.n_big); }
        bvh_big_hits(g, o, d, best, best_prim);
        if (any_ && best_prim != kInvalid) return true;
        float3 dd = d;
        if (fabsf(dd.x) < 1.0e-20f) dd.x = copysignf(1.0e-20f, dd.x);
        if (fabsf(dd.y) < 1.0e-20f) dd.y = copysignf(1.0e-20f, dd.y);
        if (fabsf(dd.z) < 1.0e-20f) dd.z = copysignf(1.0e-20f, dd.z);
        const float3 idir = f3(__fdividef(1.0f, dd.x), __fdividef(1.0f, dd.y), __fdividef(1.0f, dd.z));
        sc = f3(g.root_size[0] * (1.0f / 65535.0f) * idir.x, g.root_size[1] * (1.0f / 65535.0f) * idir.y, g.root_size[2] * (1.0f / 65535.0f) * idir.z);
        of = f3((g.root_lo[0] - o.x) * idir.x, (g.root_lo[1] - o.y) * idir.y, (g.root_lo[2] - o.z) * idir.z);
        cur = g.bvh_root;
        return cur == kBvhDone;
    }

    // 16-bit grid coordinate (low / high half of w) as a float: exact
    __device__ __forceinline__ static float q_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f; }
    __device__ __forceinline__ static float q_hi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f; }

    // entry distance of one child as an orderable key: distance bits with the slot number in the two lowest bits, or
    // 0xffffffff when the slab test fails
    __device__ __forceinline__ uint32_t child_key(float lox, float hix, float loy, float hiy, float loz, float hiz, float hi, uint32_t slot) const {
        const float ax = fmaf(lox, sc.x, of.x), bx = fmaf(hix, sc.x, of.x);
        const float ay = fmaf(loy, sc.y, of.y), by = fmaf(hiy, sc.y, of.y);
        const float az = fmaf(loz, sc.z, of.z), bz = fmaf(hiz, sc.z, of.z);
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), hi));
        return tn <= tf * 1.0000004f ? ((__float_as_uint(tn) & ~3u) | slot) : 0xffffffffu;
    }

    template <bool ALL, bool COOP> __device__ __forceinline__ bool step(const SceneAccess<ALL>& S, bool active) {
        const PathSceneD& g = *S.g;
        const int kWalkSteps = S.walk_steps;
        const uint4* __restrict__ nodes = g.bvh4_nodes;
        const bool spec = (S.bvh_spec & 1) != 0;
#pragma unroll 1
        for (int it = 0; it < kWalkSteps; ++it) {
            __syncwarp();
            if (spec && active && pend == 0u && cur >= kBvhLeaf && cur != kBvhDone) { // speculative traversal, see BvhWalk
                pend = cur;
                pop(S.stack);
            }
            if (active && cur < kBvhLeaf) {
                const uint4* __restrict__ np = nodes + 4 * (size_t)cur;
                const uint4 bx = __ldg(np), by = __ldg(np + 1), bz = __ldg(np + 2), rf = __ldg(np + 3);
                const float hi = best * 1.000002f + 1.0e-6f;
                if (COUNT) n_node += 2u + (rf.z != kBvhDone ? 1u : 0u) + (rf.w != kBvhDone ? 1u : 0u);
                // (slots 0 and 1 are always filled; an empty slot 2 / 3 carries the reference kBvhDone)
                uint32_t k0 = child_key(q_lo(bx.x), q_lo(bx.z), q_lo(by.x), q_lo(by.z), q_lo(bz.x), q_lo(bz.z), hi, 0u);
                uint32_t k1 = child_key(q_hi(bx.x), q_hi(bx.z), q_hi(by.x), q_hi(by.z), q_hi(bz.x), q_hi(bz.z), hi, 1u);
                uint32_t k2 = rf.z != kBvhDone ? child_key(q_lo(bx.y), q_lo(bx.w), q_lo(by.y), q_lo(by.w), q_lo(bz.y), q_lo(bz.w), hi, 2u) : 0xffffffffu;
                uint32_t k3 = rf.w != kBvhDone ? child_key(q_hi(bx.y), q_hi(bx.w), q_hi(by.y), q_hi(by.w), q_hi(bz.y), q_hi(bz.w), hi, 3u) : 0xffffffffu;
                // ascending order (misses last): five compare-exchanges
                uint32_t t;
                t = min(k0, k1); k1 = max(k0, k1); k0 = t;
                t = min(k2, k3); k3 = max(k2, k3); k2 = t;
                t = min(k0, k2); k2 = max(k0, k2); k0 = t;
                t = min(k1, k3); k3 = max(k1, k3); k1 = t;
                t = min(k1, k2); k2 = max(k1, k2); k1 = t;
                auto ref_of = [&](uint32_t key) {
                    const uint32_t sl = key & 3u;
                    return sl == 0u ? rf.x : (sl == 1u ? rf.y : (sl == 2u ? rf.z : rf.w));
                };
                // The hits are k0 .. k(n-1). Enter k0, postpone the others farthest first. Written without branches for
                // the common case (the capture of the branchy form ran these lines at 4-7 of 32 lanes, 15 % of the kernel's
                // instructions): three unconditional stores above the stack pointer -- what lies above it is never read.
                const int n = int(k0 != 0xffffffffu) + int(k1 != 0xffffffffu) + int(k2 != 0xffffffffu) + int(k3 != 0xffffffffu);
                const uint32_t e1 = ref_of(k1), e2 = ref_of(k2), e3 = ref_of(k3);
                if (S.bvh_spec & 2) { // A/B knob: the branchy form
                    if (n == 4) push(S.stack, e3);
                    if (n >= 3) push(S.stack, e2);
                    if (n >= 2) push(S.stack, e1);
                    if (n > 0) cur = ref_of(k0);
                    else pop(S.stack);
                    continue;
                }
                if (sp + 3 <= cap) {
                    uint32_t* __restrict__ at = S.stack + sp * kThreads;
                    at[0] = n == 4 ? e3 : (n == 3 ? e2 : e1);
                    at[kThreads] = n == 4 ? e2 : e1;
                    at[2 * kThreads] = e1;
                    sp += max(n - 1, 0);
                } else { // deep in the stack: the general form (entries beyond the shared-memory levels live in local memory)
                    if (n == 4) push(S.stack, e3);
                    if (n >= 3) push(S.stack, e2);
                    if (n >= 2) push(S.stack, e1);
                }
                if (n > 0) {
                    cur = ref_of(k0);
                } else if (sp <= cap) { // pop from the shared-memory levels: one predicated load
                    const int at = max(sp - 1, 0);
                    const uint32_t top = S.stack[at * kThreads];
                    cur = sp > 0 ? top : kBvhDone;
                    sp = at;
                } else {
                    pop(S.stack);
                }
            }
        }
        __syncwarp();
        if (active && pend == 0u && cur >= kBvhLeaf && cur != kBvhDone) {
            pend = cur;
            pop(S.stack);
        }
        if (active && pend != 0u) {
            const uint32_t first = pend & 0x0fffffffu, cnt = ((pend >> 28) & 7u) + 1u;
            pend = 0u;
            if (COUNT) n_prim += cnt;
            const float4* __restrict__ pp = g.bvh_prims + 4 * (size_t)first;
#pragma unroll 1
            for (uint32_t k = 0; k < cnt; ++k, pp += 4) {
                const float4 a0 = __ldg(pp), b0 = __ldg(pp + 1), c0 = __ldg(pp + 2), g0 = __ldg(pp + 3);
                const float t = hit_prim(a0, b0, c0, g0, o, d, best);
                if (t >= 0.0f) { best = t; best_prim = __float_as_uint(g0.w); }
            }
            if (any && best_prim != kInvalid) cur = kBvhDone;
        }
        __syncwarp();
        return active && cur == kBvhDone;
    }
};

template <int WALK> struct WalkOf { using type = TreeWalk; };
template <> struct WalkOf<1> { using type = TreeWalk2<false>; };
template <> struct WalkOf<2> { using type = TreeWalk2<true>; };
template <> struct WalkOf<3> { using type = BvhWalk; };
template <> struct WalkOf<5> { using type = Bvh4Walk<false>; };
template <> struct WalkOf<6> { using type = Bvh4Walk<true>; };
template <typename W> __device__ __forceinline__ void walk_counts(const W&, unsigned&, unsigned&) {}
template <> __device__ __forceinline__ void walk_counts(const TreeWalk2<true>& w, unsigned& nodes, unsigned& prims) {
    nodes += w.n_node;
    prims += w.n_prim;
}
template <> __device__ __forceinline__ void walk_counts(const Bvh4Walk<true>& w, unsigned& nodes, unsigned& prims) {
    nodes += w.n_node;
    prims += w.n_prim;
}
// tallies of finished rays -> totals[7] (node records visited), totals[8] (primitive tests)
__device__ __forceinline__ void flush_walk_counts(const PassArgs& a, unsigned nodes, unsigned prims) {
    nodes = warp_sum_u(nodes);
    prims = warp_sum_u(prims);
    if ((threadIdx.x & 31u) == 0u && (nodes | prims)) {
        atomicAdd(a.totals + 7, (unsigned long long)nodes);
        atomicAdd(a.totals + 8, (unsigned long long)prims);
    }
}

// Runs the walks of a whole warp to completion; MUST be called by all 32 lanes (`active` = this
// lane has a ray). Used for camera rays; secondary rays go through trace_kernel.
template <bool ALL, bool COOP, int WALK = 0>
__device__ bool traverse(const SceneAccess<ALL>& S, bool active, float3 o, float3 d, float tmax, bool any, float& t_hit,
                         uint32_t& prim_hit, unsigned* n_node = nullptr, unsigned* n_prim = nullptr) {
    typename WalkOf<WALK>::type w = {};
    w.best_prim = kInvalid;
    bool busy = active && !w.init(S, o, d, tmax, any);
    const bool walked = busy;
    while (__any_sync(kFull, busy))
        if (w.template step<ALL, COOP>(S, busy)) busy = false;
    t_hit = w.best;
    prim_hit = w.best_prim;
    if ((WALK == 2 || WALK == 6) && walked && n_node) walk_counts(w, *n_node, *n_prim);
    return active && w.best_prim != kInvalid;
}

// ---- per-slot helpers ------------------------------------------------------------
// n / d without the 32-bit integer-division sequence: float estimate + one exact correction.
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t d, uint32_t& rem) {
    uint32_t q = (uint32_t)__fdividef(__uint2float_rz(n), __uint2float_rn(d));
    int32_t r = (int32_t)(n - q * d);
    if (r < 0) { --q; r += (int32_t)d; }
    else if ((uint32_t)r >= d) { ++q; r -= (int32_t)d; }
    rem = (uint32_t)r;
    return q;
}

__device__ __forceinline__ bool slot_pixel(const PassArgs& a, uint32_t slot, int& x, int& y, uint32_t& sample) {
    uint32_t lp, tx;
    uint32_t s = fast_div(slot, a.pix_count, lp); // slot = sample in pass * window size + pixel in window
    lp += a.pix_base;
    sample = (uint32_t)a.sample_base + s;
    uint32_t lt = lp >> 10, in = lp & 1023u;
    uint32_t t = lt * (uint32_t)a.map.world + (uint32_t)a.map.rank;
    uint32_t ty = fast_div(t, (uint32_t)a.map.tiles_x, tx);
    x = int(tx * kTile + (in & 31u));
    y = int(ty * kTile + (in >> 5));
    return x < a.map.w && y < a.map.h;
}

// The reference pinhole (raytracer.h:26-30,41) with a uniform jitter inside the pixel.
__device__ __forceinline__ void camera_ray(const PassArgs& a, int x, int y, uint32_t pixel, uint32_t sample, float3& o,
                                           float3& d) {
    uint4 r = philox(pixel, sample, 0u, 0u, a.seed);
    float fx = (float(x) + u01(r.x)) * 0.0002f, fy = (float(y) + u01(r.y)) * 0.0002f;
    const PathCamera& c = a.cam;
    o = f3(c.pos[0], c.pos[1], c.pos[2]);
    float3 v = f3(c.top_left[0] - c.left[0] * fx - c.up[0] * fy, c.top_left[1] - c.left[1] * fx - c.up[1] * fy,
                  c.top_left[2] - c.left[2] * fx - c.up[2] * fy);
    d = normalize(v);
}

// Warp-ballot compaction into a queue whose space the warp reserves in private chunks of
// kChunk entries: one atomic per kChunk outputs, no block barrier. Must be called by all
// 32 lanes of a converged warp.
struct WarpCursor {
    uint32_t pos, end; // warp-uniform
    uint32_t next;     // lane 0: base of the chunk reserved ahead of time (warp_reserve)
    bool has_next;     // warp-uniform
};

__device__ __forceinline__ void warp_append(WarpCursor& c, bool want, uint32_t value, uint32_t* __restrict__ queue,
                                            uint32_t* __restrict__ counter) {
    const uint32_t m = __ballot_sync(kFull, want);
    if (m == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
    const uint32_t room = c.end - c.pos, base0 = c.pos;
    uint32_t base1 = 0;
    if (n > room) { // the tail goes to a fresh chunk
        if (lane == 0) base1 = atomicAdd(counter, kChunk);
        base1 = __shfl_sync(kFull, base1, 0);
        c.pos = base1 + (n - room);
        c.end = base1 + kChunk;
    } else {
        c.pos += n;
    }
    if (want) queue[rank < room ? base0 + rank : base1 + (rank - room)] = value;
}

// Same reservation, for records wider than a queue entry: returns the lane's position (kInvalid
// for lanes that do not want one); the caller writes the record.
// The chunk a warp will need NEXT is reserved ahead of time, as soon as an append could overflow
// the current one (room < 32): ncu attributed 5.5 % of the bounce kernel's stall samples to the 16
// instructions around this atomic -- its round trip to L2 was exposed once per 64 records. Now the
// atomic is in flight during the ~800 instructions of the next vertex and its result is only read
// at the next overflow. A warp that appends sparsely never gets close to the end of its chunk and
// prefetches nothing; a dense producer leaves at most one unused chunk behind (warp_flush pads it).
template <bool AHEAD = true, uint32_t CHUNK = kChunk>
__device__ __forceinline__ uint32_t warp_reserve(WarpCursor& c, bool want, uint32_t* __restrict__ counter) {
    static_assert(CHUNK >= 32u, "one append of a full warp must fit a fresh chunk");
    const uint32_t m = __ballot_sync(kFull, want);
    if (m == 0) return kInvalid;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
    const uint32_t room = c.end - c.pos, base0 = c.pos;
    uint32_t base1 = 0;
    if (n > room) {
        if (!c.has_next && lane == 0) c.next = atomicAdd(counter, CHUNK); // first chunk, or a sparse producer
        base1 = __shfl_sync(kFull, c.next, 0);
        c.has_next = false;
        c.pos = base1 + (n - room);
        c.end = base1 + CHUNK;
    } else {
        c.pos += n;
    }
    if (AHEAD && !c.has_next && c.end - c.pos < 32u) { // the next append may overflow: fetch its chunk now
        if (lane == 0) c.next = atomicAdd(counter, CHUNK);
        c.has_next = true;
    }
    return want ? (rank < room ? base0 + rank : base1 + (rank - room)) : kInvalid;
}

// Pad what is left of the warp's last chunk so that consumers can skip it.
__device__ __forceinline__ void warp_flush(const WarpCursor& c, uint32_t* __restrict__ queue) {
    for (uint32_t i = c.pos + (threadIdx.x & 31u); i < c.end; i += 32u) queue[i] = kInvalid;
    if (c.has_next) { // a chunk reserved ahead of time and never used
        const uint32_t base = __shfl_sync(kFull, c.next, 0);
        for (uint32_t i = threadIdx.x & 31u; i < kChunk; i += 32u) queue[base + i] = kInvalid;
    }
}

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// A slot's radiance is only ever touched by the one thread that owns the slot in a launch.
__device__ __forceinline__ void add_radiance(const PassArgs& a, uint32_t slot, float3 v) {
    // (measured: RED.ADD.F32 here costs 1.7x the whole shade kernel -- 1.2 G L2 atomics per
    // frame; the plain read-modify-write is exclusive to the slot's thread anyway)
    float* L = a.L;
    const float l0 = L[slot], l1 = L[a.plane + slot], l2 = L[2 * a.plane + slot];
    L[slot] = l0 + v.x;
    L[a.plane + slot] = l1 + v.y;
    L[2 * a.plane + slot] = l2 + v.z;
}

// The three material queues a launch feeds (the next bounce's), one warp-private cursor each.
struct Sorter {
    WarpCursor cur[3];
    uint32_t* q[3];
    uint32_t* counters; // diffuse, mirror, glass are adjacent
    uint32_t mask;      // material classes the scene has
    __device__ __forceinline__ void init(const PassArgs& a, int next_bounce) {
        const int set = (next_bounce & 1) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            cur[k] = WarpCursor{0, 0, 0, false};
            q[k] = a.q[set + k];
        }
        counters = a.counts + next_bounce * 4 + Q_DIFFUSE;
        mask = a.kind_mask;
    }
    // kind: G19_BSDF_DIFFUSE / MIRROR / GLASS, or -1 for "path ended"
    __device__ __forceinline__ void push(int kind, uint32_t slot) {
        if (mask & 1u) warp_append(cur[0], kind == G19_BSDF_DIFFUSE, slot, q[0], counters + 0);
        if (mask & 2u) warp_append(cur[1], kind == G19_BSDF_MIRROR, slot, q[1], counters + 1);
        if (mask & 4u) warp_append(cur[2], kind == G19_BSDF_GLASS, slot, q[2], counters + 2);
    }
    __device__ __forceinline__ void flush() {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (mask & (1u << k)) warp_flush(cur[k], q[k]);
    }
};

// Nearest hit of one ray, flat or tree. Called by all 32 lanes; `active` = the lane has a ray.
template <bool ALL, bool COOP = false, int WALK = 0>
__device__ __forceinline__ bool nearest(const SceneAccess<ALL>& S, bool active, float3 o, float3 d, float& t, uint32_t& prim,
                                        unsigned* n_node = nullptr, unsigned* n_prim = nullptr) {
    if constexpr (ALL) {
        bool occ;
        trace_flat<false>(S, o, d, active ? FLT_MAX : -1.0f, d, -1.0f, t, prim, occ);
        return prim != kInvalid;
    } else {
        return traverse<ALL, COOP, WALK>(S, active, o, d, FLT_MAX, false, t, prim, n_node, n_prim);
    }
}

// What a new path vertex leaves behind for the next bounce (32 B + 16 B + queue entry).
__device__ __forceinline__ void store_vertex(const PassArgs& a, uint32_t slot, float3 p, uint32_t prim, float3 d,
                                             uint32_t pixel) {
    a.hp[slot] = make_float4(p.x, p.y, p.z, __uint_as_float(prim));
    a.dw[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(pixel));
}

// ---- raygen + extend (camera segment), tree scenes: slot-indexed state ----------------------
template <int OCC, bool COOP, int WALK> __global__ void __launch_bounds__(kThreads, OCC) raygen_extend_kernel(const PassArgs a) {
    pdl_launch_dependents();
    const uint32_t n = a.n_slots;
    if (blockIdx.x * kThreads >= n) return;
    const SceneAccess<false> S = stage_scene<false>(a);
    pdl_wait(); // the previous pass's accumulate cleared the queue lengths and the radiance planes
    Sorter out;
    out.init(a, 0);
    unsigned cnt_node = 0, cnt_prim = 0;
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q - lane < n; q += stride) { // warp-uniform trip count
        const uint32_t slot = q;
        int kind = -1;
        bool live = false;
        uint32_t pixel = 0;
        float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
        if (q < n) {
            int x, y;
            uint32_t sample;
            live = slot_pixel(a, slot, x, y, sample);
            if (live) {
                pixel = uint32_t(y) * uint32_t(a.map.w) + uint32_t(x);
                camera_ray(a, x, y, pixel, sample, o, d);
            }
        }
        float t;
        uint32_t prim;
        if (nearest<false, COOP, WALK>(S, live, o, d, t, prim, &cnt_node, &cnt_prim)) { // all 32 lanes walk together
            const float4 tag = S.hot_row(prim, 3);
            const int bsdf = __float_as_int(tag.y); // material class rides in the hot record
            if (bsdf == G19_BSDF_EMITTER) {          // directly visible light: the path ends here
                const MaterialD& m = a.scene.materials[__float_as_int(tag.x)];
                add_radiance(a, slot, f3(m.emission[0], m.emission[1], m.emission[2]));
            } else if (bsdf == G19_BSDF_DIFFUSE || a.max_depth > 1) {
                kind = bsdf;
                store_vertex(a, slot, o + d * t, prim, d, pixel);
            }
        }
        out.push(kind, slot);
    }
    out.flush();
    if (WALK == 2 || WALK == 6) flush_walk_counts(a, cnt_node, cnt_prim);
}

// ---- bounce: shade + trace the continuation ray ------------------------------------------
__device__ __forceinline__ void onb(float3 n, float3& t, float3& b) { // Duff et al. 2017
    float s = copysignf(1.0f, n.z);
    float a = -1.0f / (s + n.z);
    float bb = n.x * n.y * a;
    t = f3(1.0f + s * n.x * n.x * a, s * bb, -s * n.x);
    b = f3(bb, s + n.y * n.y * a, -n.y);
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- shading of one path vertex (shared by the flat-scene and the tree-scene bounce kernels) ----
// In: the vertex (hit point p on primitive prim, arrival direction d, throughput T). Out: the
// next-event sample (direction w, length tmax_s, radiance lit_rgb if it turns out unoccluded) and
// the continuation ray (origin no, direction nd, throughput T updated). flags bit 0: specular.
struct Shaded {
    float3 no, nd, w, lit_rgb, T;
    float tmax_s;
    bool want_shadow;
    uint32_t flags;
};
template <int KIND, bool LAST, bool ALL>
__device__ __forceinline__ Shaded shade_vertex(const PassArgs& a, const SceneAccess<ALL>& S, int bounce, float3 p, float3 d,
                                               float3 T, uint32_t prim, uint32_t pixel, uint32_t sample) {
    constexpr bool kDiffuse = KIND == Q_DIFFUSE;
    Shaded o;
    float4 c0, c1; // (normal, ior), (albedo, material)
    S.cold(prim, c0, c1);
    float3 ng = f3(c0.x, c0.y, c0.z);
    if (c0.x == 0.0f && c0.y == 0.0f && c0.z == 0.0f) { // spheres store no normal: (p - centre) / r
        const float4 q0 = S.hot_row(prim, 0);
        ng = (p - f3(q0.x, q0.y, q0.z)) * __fdividef(1.0f, q0.w);
    }
    const float ior = c0.w;
    const bool entering = dot(ng, d) < 0.0f;
    const float3 nf = entering ? ng : -ng; // normal on the side the ray arrives from
    const float3 albedo = f3(c1.x, c1.y, c1.z);
    const uint4 r = philox(pixel, sample, uint32_t(bounce), 1u, a.seed);
    o.nd = f3(0.f, 0.f, 1.f);
    o.flags = 0;
    o.want_shadow = false;
    o.w = f3(0.f, 0.f, 1.f);
    o.lit_rgb = f3(0.f, 0.f, 0.f);
    o.tmax_s = -1.0f;
    if (kDiffuse) {
        // next-event estimation (diffuse only): one light, one uniformly sampled point
        if (a.scene.n_lights > 0) {
            float pick = u01(r.x) * float(a.scene.n_lights);
            int li = min(int(pick), a.scene.n_lights - 1);
            float u1 = pick - float(li), u2 = u01(r.y);
            const float4* lt = S.light(li);
            const float4 l0 = lt[0], l1 = lt[1], l2 = lt[2], l3 = lt[3];
            float su = sqrtf(u1);
            float b1 = su * (1.0f - u2), b2 = su * u2;
            float3 yl = f3(l0.x + l1.x * b1 + l2.x * b2, l0.y + l1.y * b1 + l2.y * b2, l0.z + l1.z * b1 + l2.z * b2);
            float3 w = yl - p;
            float dist2 = dot(w, w);
            float inv_dist = rsqrtf(dist2);
            float dist = dist2 * inv_dist;
            w = w * inv_dist;
            o.w = w;
            float cs = dot(nf, w);
            float cl = fabsf(dot(f3(l3.x, l3.y, l3.z), w));
            if (cs > 0.0f && cl > 0.0f && dist > 2.0f * kRayEps) {
                o.want_shadow = true;
                o.tmax_s = dist - 2.0f * kRayEps;
                const float4 l4 = lt[4];
                float gterm = cs * cl * l0.w * __fdividef(1.0f, dist2 * l1.w) * (1.0f / kPi);
                o.lit_rgb = f3(T.x * albedo.x * l4.x * gterm, T.y * albedo.y * l4.y * gterm, T.z * albedo.z * l4.z * gterm);
            }
        }
        o.no = p + nf * kRayEps;
        if (!LAST) { // cosine-weighted bounce: pdf cancels cos/pi, throughput *= albedo
            float u3 = u01(r.z), u4 = u01(r.w);
            float rr = sqrtf(u3), phi = 2.0f * kPi * u4 - kPi; // [-pi, pi): MUFU range
            float sp, cp;
            __sincosf(phi, &sp, &cp);
            sp = -sp; cp = -cp; // shift back by pi
            float3 tx, ty;
            onb(nf, tx, ty);
            o.nd = normalize(tx * (rr * cp) + ty * (rr * sp) + nf * sqrtf(fmaxf(0.0f, 1.0f - u3)));
        }
        o.T = T * albedo;
    } else if (KIND == Q_MIRROR) {
        o.nd = normalize(d - nf * (2.0f * dot(d, nf)));
        o.no = p + nf * kRayEps;
        o.T = T * albedo;
        o.flags = 1u;
    } else { // dielectric
        float etai = entering ? 1.0f : ior, etat = entering ? ior : 1.0f;
        float eta = etai / etat;
        float cosi = fminf(1.0f, -dot(d, nf));
        float sin2t = eta * eta * fmaxf(0.0f, 1.0f - cosi * cosi);
        float F = 1.0f;
        float cost = 0.0f;
        if (sin2t < 1.0f) {
            cost = sqrtf(1.0f - sin2t);
            float rs = (etai * cosi - etat * cost) / (etai * cosi + etat * cost);
            float rp = (etai * cost - etat * cosi) / (etai * cost + etat * cosi);
            F = 0.5f * (rs * rs + rp * rp);
        }
        if (u01(r.x) < F) {
            o.nd = normalize(d + nf * (2.0f * cosi));
            o.no = p + nf * kRayEps;
        } else {
            o.nd = normalize(d * eta + nf * (eta * cosi - cost));
            o.no = p - nf * kRayEps;
        }
        o.T = T * albedo;
        o.flags = 1u;
    }
    return o;
}

// ---- flat scenes: dense vertex records --------------------------------------------------
// What the ncu capture of the slot-indexed state showed (profiles/r01c_ncu_pass_summary.json): the four
// diffuse bounce launches of a Cornell pass execute 100 / 93 / 83 / 73 M warp instructions but all
// take ~128 us -- as paths end, the survivors' slots thin out, every 16-byte state record still
// costs a 32-byte sector (global load sectors per launch stay at 8 M), DRAM channel load becomes
// uneven and long-scoreboard stalls grow 0.55 -> 1.97 per issue. So the QUEUES now hold the records
// themselves: a vertex is written at the position its warp reserved in the next bounce's material
// queue (slot, hit point + primitive, direction + pixel, throughput + sample, radiance so far --
// 64 bytes, all planes coalesced) and read back from consecutive addresses; the path's radiance
// travels with it and is stored to the slot's accumulator input exactly once, when the path ends.
// A bounce holds at most P vertices over its three material queues together, so the queues of a bounce parity share
// storage instead of being P entries each (16 GB reserved for a 14-primitive scene): the diffuse queue has
// an array of its own (P + slack entries), mirror and glass share ONE array of P + 2 x slack entries -- mirror grows
// from its bottom, glass from its top downwards (entry q of the glass queue lives at top - q: consecutive lanes still
// touch consecutive 16-byte words). A diffuse-only scene allocates no specular array at all.
struct RecView { // one queue: four float4 planes; entry q is at plane[q * dir]
    float4 *ls, *hp, *dw, *tp; // (radiance so far | slot) (hit point | primitive) (direction | pixel) (throughput | sample)
    int dir;
    __device__ __forceinline__ ptrdiff_t at(uint32_t q) const { return dir > 0 ? ptrdiff_t(q) : -ptrdiff_t(q); }
};
__device__ __forceinline__ RecView rec_queue(const PassArgs& a, int qi) { // qi = bounce parity * 3 + (kind - 1)
    RecView r;
    const int parity = qi >= 3 ? 1 : 0, kind = qi - 3 * parity;
    size_t off = size_t(parity) * (a.queue_cap + a.spec_cap);
    r.dir = 1;
    if (kind != 0) {
        off += a.queue_cap;                                      // the shared specular array of this parity
        if (kind == 2) { off += a.spec_cap - 1; r.dir = -1; }   // glass: from the top, downwards
    }
    r.ls = a.rec_ls + off;
    r.hp = a.rec_hp + off;
    r.dw = a.rec_dw + off;
    r.tp = a.rec_tp + off;
    return r;
}

// The record queues a launch feeds (the next bounce's). SPEC = the scene has mirror or glass
// materials; a diffuse-only scene keeps one cursor (registers: the three-cursor form spilled).
template <bool SPEC> struct RecSorter {
    WarpCursor cur[SPEC ? 3 : 1];
    uint32_t* counters;
    uint32_t mask;
    int set;
    __device__ __forceinline__ void init(const PassArgs& a, int next_bounce) {
#pragma unroll
        for (int k = 0; k < (SPEC ? 3 : 1); ++k) cur[k] = WarpCursor{0, 0, 0, false};
        set = (next_bounce & 1) * 3;
        counters = a.counts + next_bounce * 4 + Q_DIFFUSE;
        mask = SPEC ? a.kind_mask : 1u;
    }
    // kind: G19_BSDF_DIFFUSE / MIRROR / GLASS, or -1 for "path ended". Returns the lane's position in
    // its kind's queue (kInvalid for ended paths). Called by all 32 lanes.
    __device__ __forceinline__ uint32_t reserve(int kind) {
        if (!SPEC) return warp_reserve(cur[0], kind == 0, counters);
        uint32_t pos = kInvalid;
#pragma unroll
        for (int k = 0; k < (SPEC ? 3 : 1); ++k)
            if (mask & (1u << k)) {
                // the specular queues fill slowly (a few entries per warp and launch): no chunk ahead of
                // time and 32-entry chunks, or their consumers would run on mostly padding
                const uint32_t p = k == 0 ? warp_reserve<true>(cur[k], kind == k, counters + k)
                                          : warp_reserve<false, 32u>(cur[k], kind == k, counters + k);
                if (kind == k) pos = p;
            }
        return pos;
    }
    // pad what is left of the warp's chunks: an entry whose slot word is kInvalid is skipped by the consumer
    __device__ __forceinline__ void flush(const PassArgs& a) {
        const uint32_t lane = threadIdx.x & 31u;
        const float4 pad = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInvalid));
#pragma unroll
        for (int k = 0; k < (SPEC ? 3 : 1); ++k)
            if (mask & (1u << k)) {
                const RecView r = rec_queue(a, set + k);
                const WarpCursor& c = cur[k];
                for (uint32_t i = c.pos + lane; i < c.end; i += 32u) r.ls[r.at(i)] = pad;
                if (c.has_next) { // a chunk reserved ahead of time and never used
                    const uint32_t base = __shfl_sync(kFull, c.next, 0);
                    for (uint32_t i = lane; i < kChunk; i += 32u) r.ls[r.at(base + i)] = pad;
                }
            }
    }
};

// ---- raygen + extend, flat scenes ------------------------------------------------------
template <bool SPEC> __global__ void __launch_bounds__(kThreads, 4) raygen_extend_flat_kernel(const PassArgs a) {
    pdl_launch_dependents();
    const uint32_t n = a.n_slots;
    if (blockIdx.x * kThreads >= n) return;
    const SceneAccess<true> S = stage_scene<true>(a);
    pdl_wait(); // the previous pass's accumulate cleared the queue lengths and the radiance planes
    RecSorter<SPEC> out;
    out.init(a, 0);
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q - lane < n; q += stride) { // warp-uniform trip count
        const uint32_t slot = q;
        int kind = -1;
        bool live = false;
        uint32_t pixel = 0, prim = kInvalid, sample = 0;
        float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
        float t = 0.0f;
        if (q < n) {
            int x, y;
            live = slot_pixel(a, slot, x, y, sample);
            if (live) {
                pixel = uint32_t(y) * uint32_t(a.map.w) + uint32_t(x);
                camera_ray(a, x, y, pixel, sample, o, d);
            }
        }
        float4 seen = make_float4(0.f, 0.f, 0.f, 0.f); // what a path that ends on the camera segment delivers
        if (nearest<true>(S, live, o, d, t, prim)) {
            const float4 tag = S.hot_row(prim, 3);
            const int bsdf = __float_as_int(tag.y); // material class rides in the hot record
            if (bsdf == G19_BSDF_EMITTER) {          // directly visible light: the path ends here
                const MaterialD& m = a.scene.materials[__float_as_int(tag.x)];
                seen = make_float4(m.emission[0], m.emission[1], m.emission[2], 0.0f);
            } else if (bsdf == G19_BSDF_DIFFUSE || a.max_depth > 1) {
                kind = bsdf;
            }
        }
        // Every slot of the pass gets its radiance stored exactly once, by whichever kernel ends its
        // path (here: miss, emitter, or a slot outside the frame), so `accumulate` only reads.
        if (q < n && kind < 0) reinterpret_cast<float4*>(a.L)[slot] = seen;
        const uint32_t pos = out.reserve(kind);
        if (pos != kInvalid) { // throughput is 1 and the radiance 0 on the camera segment: not stored
            const RecView r = rec_queue(a, kind);
            const float3 p = o + d * t;
            const ptrdiff_t at = r.at(pos);
            r.ls[at] = make_float4(__uint_as_float(sample), 0.f, 0.f, __uint_as_float(slot)); // radiance is 0 here: x carries the sample index
            r.hp[at] = make_float4(p.x, p.y, p.z, __uint_as_float(prim));
            r.dw[at] = make_float4(d.x, d.y, d.z, __uint_as_float(pixel));
        }
    }
    out.flush(a);
}

// Asynchronous prefetch of the NEXT record into shared memory (LDGSTS): a register prefetch gets
// sunk to its first use by the compiler (ncu: 20 % of the kernel's stall samples sat on that one
// instruction); a cp.async cannot be, and it holds no registers while in flight. Two buffers per
// thread, each thread only ever touches its own column. The records are dense, so the address of
// the next one is known without loading anything first.
template <bool FULL> struct RecStage {
    float4 ls[2][kThreads]; // radiance so far, slot
    float4 hp[2][kThreads];
    float4 dw[2][kThreads];
    float4 tp[FULL ? 2 : 1][FULL ? kThreads : 1];
};
template <bool FULL>
__device__ __forceinline__ void prefetch_rec(const RecView& r, uint32_t pos, bool valid, RecStage<FULL>& st, int buf) {
    const int t = threadIdx.x;
    if (valid) {
        const ptrdiff_t at = r.at(pos);
        cp_async16(&st.ls[buf][t], r.ls + at);
        cp_async16(&st.hp[buf][t], r.hp + at);
        cp_async16(&st.dw[buf][t], r.dw + at);
        if (FULL) cp_async16(&st.tp[buf][t], r.tp + at);
    }
    cp_async_commit();
}

// KIND: Q_DIFFUSE / Q_MIRROR / Q_GLASS. FIRST: the vertex of the camera segment. LAST: the path's
// final segment ended here -- next-event estimation only, no continuation.
// One launch per material queue and bounce: next-event estimation, BSDF sample, the trace of BOTH
// rays in one loop over the staged primitives, and the new vertex record written straight into
// the next bounce's material queue.
// FUSED (flat scenes, first bounce): the camera segment is traced right here instead of by
// raygen_extend_flat_kernel -- the vertex record of the camera hit (48 B written, 48 B read back, a fifth of the frame's
// HBM traffic), its queue reservation and a launch per pass disappear; `n` is then the number of path slots of the pass.
// FOLD (flat scenes, the bounce before the last): the vertex the continuation ray finds would be the path's
// last one -- next-event estimation only. It is shaded right here (light sample + shadow ray) instead of being written as a
// 64-byte record for a launch of its own to read back: same arithmetic in the same order, one launch per pass fewer.
template <int KIND, bool FIRST, bool LAST, bool SPEC, bool SPLIT = false, bool FUSED = false, bool FOLD = false>
__device__ __forceinline__ void bounce_flat_body(const PassArgs& a, const int bounce, const SceneAccess<true>& S, const uint32_t n,
                                                 const uint32_t cta, const uint32_t n_cta, RecStage<!FIRST>& stage) {
    constexpr bool kDiffuse = KIND == Q_DIFFUSE;
    const RecView in = rec_queue(a, (bounce & 1) * 3 + (KIND - 1));
    RecSorter<SPEC> out;
    if (!LAST && !FOLD) out.init(a, bounce + 1);
    // statistics (g19_stats), two 16-bit counters per register (a thread runs < 2^16 iterations: the grid
    // has > 10^5 threads and a queue < 2^32 entries): calls | traced << 16, shadow rays | lit << 16
    uint32_t cnt_a = 0, cnt_b = 0, stored = 0, folded = 0;
    const uint32_t stride = n_cta * kThreads;
    const uint32_t lane = threadIdx.x & 31u;
    const int tid = threadIdx.x;
    const bool next_last = bounce + 2 >= a.max_depth;

    uint32_t q = cta * kThreads + threadIdx.x;
    int buf = 0;
    if (!FUSED) prefetch_rec(in, q, q < n, stage, 0);
    for (; q - lane < n; q += stride) {
        uint32_t slot = kInvalid;
        float4 ls = make_float4(0.f, 0.f, 0.f, 0.f);
        float3 p = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
        uint32_t prim = kInvalid, pixel = 0, sample = 0;
        if (FUSED) {
            // ---- the camera segment of path slot q (what raygen_extend_flat_kernel does for scenes with specular materials) ----
            bool live = false;
            float3 o = f3(0.f, 0.f, 0.f);
            if (q < n) {
                int x, y;
                live = slot_pixel(a, q, x, y, sample);
                if (live) {
                    pixel = uint32_t(y) * uint32_t(a.map.w) + uint32_t(x);
                    camera_ray(a, x, y, pixel, sample, o, d);
                }
            }
            float t = 0.0f;
            uint32_t pr = kInvalid;
            int spec_kind = -1; // SPEC: the camera ray met a mirror / glass surface: that vertex goes to its material's queue
            float4 seen = make_float4(0.f, 0.f, 0.f, 0.f); // what a path that ends on the camera segment delivers
            if (nearest<true>(S, live, o, d, t, pr)) {
                const float4 tag = S.hot_row(pr, 3);
                const int bsdf = __float_as_int(tag.y);
                if (bsdf == G19_BSDF_EMITTER) { // directly visible light: the path ends here
                    const MaterialD& m = a.scene.materials[__float_as_int(tag.x)];
                    seen = make_float4(m.emission[0], m.emission[1], m.emission[2], 0.0f);
                } else if (bsdf == G19_BSDF_DIFFUSE) {
                    slot = q;
                    prim = pr;
                    p = o + d * t;
                } else if (SPEC && a.max_depth > 1) {
                    spec_kind = bsdf;
                }
            }
            if (q < n && slot == kInvalid && spec_kind < 0) reinterpret_cast<float4*>(a.L)[q] = seen; // miss, emitter, or a slot outside the frame
            if (SPEC) {
                // The few camera hits on mirror / glass (the two spheres of the glass Cornell box: a tenth of the pixels) keep
                // the material queues: their 48-byte camera records go to bounce 0's queue of their kind, reserved with one
                // atomic per warp and kind -- no cursor state carried through the loop -- for the launch that follows this one.
#pragma unroll
                for (int k = G19_BSDF_MIRROR; k <= G19_BSDF_GLASS; ++k) {
                    const uint32_t m = __ballot_sync(kFull, spec_kind == k);
                    if (m == 0u) continue;
                    uint32_t base = 0;
                    if (lane == uint32_t(__ffs(int(m)) - 1)) base = atomicAdd(a.counts + Q_DIFFUSE + k, uint32_t(__popc(m)));
                    base = __shfl_sync(kFull, base, __ffs(int(m)) - 1);
                    if (spec_kind == k) {
                        const RecView r = rec_queue(a, k); // bounce 0, parity 0
                        const ptrdiff_t at = r.at(base + uint32_t(__popc(m & ((1u << lane) - 1u))));
                        const float3 ph = o + d * t;
                        r.ls[at] = make_float4(__uint_as_float(sample), 0.f, 0.f, __uint_as_float(q));
                        r.hp[at] = make_float4(ph.x, ph.y, ph.z, __uint_as_float(pr));
                        r.dw[at] = make_float4(d.x, d.y, d.z, __uint_as_float(pixel));
                    }
                }
            }
        } else {
            const uint32_t qn = q + stride;
            prefetch_rec(in, qn, qn < n && qn > q, stage, buf ^ 1); // next record, in flight during this body
            cp_async_wait<1>();                                     // this record has landed
            ls = stage.ls[buf][tid];
            slot = q < n ? __float_as_uint(ls.w) : kInvalid;
        }
        int kind_next = -1;
        float3 p_next = f3(0.f, 0.f, 0.f), d_next = f3(0.f, 0.f, 1.f), T = f3(1.f, 1.f, 1.f), Lp = f3(0.f, 0.f, 0.f);
        uint32_t prim_next = kInvalid;
        if (slot != kInvalid) {
            if (!FUSED) {
                const float4 hp = stage.hp[buf][tid], dw = stage.dw[buf][tid];
                p = f3(hp.x, hp.y, hp.z);
                d = f3(dw.x, dw.y, dw.z);
                prim = __float_as_uint(hp.w);
                pixel = __float_as_uint(dw.w);
            }
            if (FUSED) {
                // throughput 1, no radiance yet: nothing to load
            } else if (FIRST) {
                sample = __float_as_uint(ls.x); // the camera record: no radiance yet, x carries the sample index
            } else {
                const float4 tp = stage.tp[buf][tid];
                T = f3(tp.x, tp.y, tp.z);
                sample = __float_as_uint(tp.w);
                Lp = f3(ls.x, ls.y, ls.z);
            }
            const Shaded sh = shade_vertex<KIND, LAST, true>(a, S, bounce, p, d, T, prim, pixel, sample);
            T = sh.T;
            const bool cont = !LAST && (T.x > 0.0f || T.y > 0.0f || T.z > 0.0f);
            cnt_a += 1u + (cont ? 0x10000u : 0u);
            // ---- the rays of this vertex: shadow (any hit) and continuation (nearest hit) ----
            bool blocked = false;
            float t_hit = FLT_MAX;
            uint32_t prim_hit = kInvalid;
            if (kDiffuse && !LAST && SPLIT) {
                // (bounce_occ = 4 experiment: one ray at a time through the primitive loops -- fewer live registers, the
                // origin half of the affine maps computed twice)
                if (sh.want_shadow) {
                    bool unused;
                    float ts;
                    uint32_t ps;
                    trace_flat<false>(S, sh.no, sh.w, sh.tmax_s, sh.w, -1.0f, ts, ps, unused);
                    blocked = ps != kInvalid;
                }
                if (cont) {
                    bool unused;
                    trace_flat<false>(S, sh.no, sh.nd, FLT_MAX, sh.nd, -1.0f, t_hit, prim_hit, unused);
                }
            } else if (kDiffuse && !LAST) {
                if (sh.want_shadow || cont)
                    trace_flat<true>(S, sh.no, sh.nd, cont ? FLT_MAX : -1.0f, sh.w, sh.tmax_s, t_hit, prim_hit, blocked);
            } else if (kDiffuse) { // LAST: the shadow ray alone
                if (sh.want_shadow) {
                    bool unused;
                    trace_flat<false>(S, sh.no, sh.w, sh.tmax_s, sh.w, -1.0f, t_hit, prim_hit, unused);
                    blocked = prim_hit != kInvalid;
                }
            } else if (cont) {
                bool unused;
                trace_flat<false>(S, sh.no, sh.nd, FLT_MAX, sh.nd, -1.0f, t_hit, prim_hit, unused);
            }
            const bool hit = cont && prim_hit != kInvalid;
            const bool lit = kDiffuse && sh.want_shadow && !blocked;
            if (kDiffuse) cnt_b += (sh.want_shadow ? 1u : 0u) + (lit ? 0x10000u : 0u);
            if (lit) Lp = Lp + sh.lit_rgb;
            if (hit) {
                const float4 tag = S.hot_row(prim_hit, 3);
                const int bsdf = __float_as_int(tag.y);
                if (bsdf == G19_BSDF_EMITTER) {
                    // emission counts after specular bounces only (NEE covers the diffuse ones)
                    if (sh.flags & 1u) {
                        const MaterialD& m = a.scene.materials[__float_as_int(tag.x)];
                        Lp = Lp + f3(T.x * m.emission[0], T.y * m.emission[1], T.z * m.emission[2]);
                    }
                } else if (bsdf == G19_BSDF_DIFFUSE || !next_last) { // a specular vertex on the last segment adds nothing
                    kind_next = bsdf;
                    p_next = sh.no + sh.nd * t_hit;
                    d_next = sh.nd;
                    prim_next = prim_hit;
                }
            }
            if (FOLD && kind_next == G19_BSDF_DIFFUSE) { // the path's last vertex, shaded in place (next_last holds for this launch)
                const Shaded s2 = shade_vertex<Q_DIFFUSE, true, true>(a, S, bounce + 1, p_next, d_next, T, prim_next, pixel, sample);
                cnt_a += 1u;
                ++folded;
                if (s2.want_shadow) {
                    bool unused;
                    float t2;
                    uint32_t k2;
                    trace_flat<false>(S, s2.no, s2.w, s2.tmax_s, s2.w, -1.0f, t2, k2, unused);
                    const bool lit2 = k2 == kInvalid;
                    cnt_b += 1u + (lit2 ? 0x10000u : 0u);
                    if (lit2) Lp = Lp + s2.lit_rgb;
                }
                kind_next = -1;
            }
            if (kind_next < 0) {
                // the path ends here: its radiance goes to the slot's accumulator input, once (zero or not:
                // accumulate does not clear what it read)
                ++stored;
                // (slots are scattered by now: ONE 16-byte store, not three 4-byte stores into three planes --
                // the last bounce of the depth-12 glass box spent 148 us on 37 M instructions doing those)
                reinterpret_cast<float4*>(a.L)[slot] = make_float4(Lp.x, Lp.y, Lp.z, 0.0f);
            }
        }
        if (!LAST && !FOLD) {
            const uint32_t pos = out.reserve(kind_next);
            if (pos != kInvalid) {
                const RecView r = rec_queue(a, out.set + kind_next);
                const ptrdiff_t at = r.at(pos);
                r.ls[at] = make_float4(Lp.x, Lp.y, Lp.z, __uint_as_float(slot));
                r.hp[at] = make_float4(p_next.x, p_next.y, p_next.z, __uint_as_float(prim_next));
                r.dw[at] = make_float4(d_next.x, d_next.y, d_next.z, __uint_as_float(pixel));
                r.tp[at] = make_float4(T.x, T.y, T.z, __uint_as_float(sample));
            }
        }
        buf ^= 1;
    }
    cp_async_wait<0>();
    if (!LAST && !FOLD) out.flush(a);
    const unsigned calls = warp_sum(cnt_a & 0xffffu), traced = warp_sum(cnt_a >> 16);
    const unsigned shadow_rays = warp_sum(cnt_b & 0xffffu), lit = warp_sum(cnt_b >> 16);
    stored = warp_sum(stored);
    if (lane == 0 && calls) {
        atomicAdd(a.totals + 2, (unsigned long long)calls);
        if (FIRST) atomicAdd(a.totals + 3, (unsigned long long)calls);
        if (traced) atomicAdd(a.totals + 0, (unsigned long long)traced);
        if (shadow_rays) atomicAdd(a.totals + 1, (unsigned long long)shadow_rays);
        if (lit) atomicAdd(a.totals + 4, (unsigned long long)lit);
        if (stored) atomicAdd(a.totals + 6, (unsigned long long)stored);
    }
    if (FOLD) {
        folded = warp_sum(folded);
        if (lane == 0 && folded) atomicAdd(a.totals + 9, (unsigned long long)folded);
    }
}

// One material queue per launch (diffuse-only scenes, and the last bounce of any scene).
template <int KIND, bool FIRST, bool LAST, bool SPEC, int OCC = 3, bool FOLD = false>
__global__ void __launch_bounds__(kThreads, KIND != Q_DIFFUSE ? 4 : OCC) bounce_flat_kernel(const PassArgs a, const int bounce) {
    pdl_launch_dependents();
    if (LAST && KIND != Q_DIFFUSE) return; // a specular vertex on the last segment contributes nothing
    const SceneAccess<true> S = stage_scene<true>(a); // does not depend on earlier kernels: overlaps their tail
    pdl_wait();
    const uint32_t n = a.counts[bounce * 4 + KIND];
    if (blockIdx.x * kThreads >= n) return; // short queue: surplus CTAs leave
    __shared__ RecStage<!FIRST> stage;
    bounce_flat_body<KIND, FIRST, LAST, SPEC, OCC == 4, false, FOLD>(a, bounce, S, n, blockIdx.x, gridDim.x, stage);
}

// Diffuse-only flat scenes: camera segment + first vertex in one launch (bounce_flat_body<FUSED>), over the path slots of
// the pass instead of a queue.
template <bool LAST, bool FOLD = false, bool SPEC = false> __global__ void __launch_bounds__(kThreads, 3) bounce_flat_fused_kernel(const PassArgs a) {
    pdl_launch_dependents();
    const uint32_t n = a.n_slots;
    if (blockIdx.x * kThreads >= n) return;
    const SceneAccess<true> S = stage_scene<true>(a);
    pdl_wait(); // the previous pass's accumulate cleared the queue lengths
    RecStage<false>& unused = *reinterpret_cast<RecStage<false>*>(g19_dyn_smem); // (the fused body prefetches no records)
    bounce_flat_body<Q_DIFFUSE, true, LAST, SPEC, false, true, FOLD>(a, 0, S, n, blockIdx.x, gridDim.x, unused);
}

// All three material queues of one bounce in ONE launch (scenes with mirror / glass). The mirror and
// glass queues of a Cornell box hold a few per cent of the vertices: as launches of their own they
// ran 1-2 records per thread at 50-60 % issue utilisation with nothing to overlap their latency
// (ncu, profiles/r01c: 25 + 35 us next to the diffuse launch's 216 us, 21 of 34 launches per pass).
// Here the grid's CTAs are split between the queues in proportion to their estimated work, so the
// specular vertices fill issue slots beside the diffuse ones and a pass is 13 launches, not 34.
// The per-material queues and kernels' code are unchanged: a CTA runs exactly one KIND.
template <bool FIRST, bool FOLD = false>
__global__ void __launch_bounds__(kThreads, 3) bounce_flat_all_kernel(const PassArgs a, const int bounce) {
    pdl_launch_dependents();
    const SceneAccess<true> S = stage_scene<true>(a);
    pdl_wait();
    const uint32_t n1 = a.counts[bounce * 4 + Q_DIFFUSE], n2 = a.counts[bounce * 4 + Q_MIRROR], n3 = a.counts[bounce * 4 + Q_GLASS];
    // estimated work per queue entry relative to a diffuse vertex (two rays, light sample): measured
    // instructions per entry, padding of the sparsely filled specular chunks included
    const float w1 = float(n1), w2 = 0.45f * float(n2), w3 = 0.6f * float(n3);
    const float total = w1 + w2 + w3;
    if (total <= 0.0f) return;
    const uint32_t G = gridDim.x;
    uint32_t g2 = n2 ? max(1u, min(uint32_t(float(G) * w2 / total + 0.5f), (n2 + kThreads - 1) / kThreads)) : 0u;
    uint32_t g3 = n3 ? max(1u, min(uint32_t(float(G) * w3 / total + 0.5f), (n3 + kThreads - 1) / kThreads)) : 0u;
    if (g2 + g3 >= G) { g2 = min(g2, G / 3); g3 = min(g3, G / 3); }
    const uint32_t g1 = n1 ? min(G - g2 - g3, (n1 + kThreads - 1) / kThreads) : 0u;
    __shared__ RecStage<!FIRST> stage;
    const uint32_t b = blockIdx.x;
    if (b < g1) bounce_flat_body<Q_DIFFUSE, FIRST, false, true, false, false, FOLD>(a, bounce, S, n1, b, g1, stage);
    else if (b < g1 + g2) bounce_flat_body<Q_MIRROR, FIRST, false, true, false, false, FOLD>(a, bounce, S, n2, b - g1, g2, stage);
    else if (b < g1 + g2 + g3) bounce_flat_body<Q_GLASS, FIRST, false, true, false, false, FOLD>(a, bounce, S, n3, b - g1 - g2, g3, stage);
}

// ---- bounce, tree scenes: shade and queue the rays -----------------------------------------
// Asynchronous prefetch of the NEXT slot's state into shared memory (see prefetch_rec); here the
// state is indexed by slot, so the queue entry is loaded two iterations ahead.
template <bool TP> struct VertexStage {
    float4 hp[2][kThreads];
    float4 dw[2][kThreads];
    float4 tp[TP ? 2 : 1][TP ? kThreads : 1];
};
template <bool TP> __device__ __forceinline__ void prefetch_vertex(const PassArgs& a, uint32_t slot, VertexStage<TP>& st, int buf) {
    const int t = threadIdx.x;
    if (slot != kInvalid) {
        cp_async16(&st.hp[buf][t], a.hp + slot);
        cp_async16(&st.dw[buf][t], a.dw + slot);
        if (TP) cp_async16(&st.tp[buf][t], a.tp + slot);
    }
    cp_async_commit();
}

// Tree scenes only shade here: the rays go to the bounce's ray queue and trace_kernel walks them
// with dynamic fetch (a heavy tail of long walks would otherwise hold the other 31 lanes of the
// warp -- ncu: 3.4 of 32 lanes active); trace_kernel also adds the light sample and delivers the
// continuation vertex into the slot-indexed state and the next bounce's material queues.
// Sort key of a queued ray (ray_sort.cu): where it starts, on a 32^3 grid over the root box, in Morton order.
// Shadow rays (they all aim at the lights): 1 | 15 bits; continuation rays: 0 | 12 bits (16^3 cells) | direction octant.
__device__ __forceinline__ uint32_t spread5(uint32_t x) { // bit k -> bit 3k
    x = (x | (x << 8)) & 0x100fu;
    x = (x | (x << 4)) & 0x10c3u;
    x = (x | (x << 2)) & 0x1249u;
    return x;
}
__device__ __forceinline__ uint32_t ray_sort_key(const PathSceneD& g, float3 o, float3 d, bool shadow) {
    const int sh = kMaxTreeDepth - 5;
    const uint32_t x = uint32_t(min(max(__float2int_rd((o.x - g.root_lo[0]) * g.grid_scale[0]) >> sh, 0), 31));
    const uint32_t y = uint32_t(min(max(__float2int_rd((o.y - g.root_lo[1]) * g.grid_scale[1]) >> sh, 0), 31));
    const uint32_t z = uint32_t(min(max(__float2int_rd((o.z - g.root_lo[2]) * g.grid_scale[2]) >> sh, 0), 31));
    const uint32_t m = spread5(x) | (spread5(y) << 1) | (spread5(z) << 2);
    if (shadow) return 0x8000u | m;
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    return ((m >> 3) << 3) | oct;
}

template <int KIND, bool FIRST, bool LAST>
__global__ void __launch_bounds__(kThreads, KIND != Q_DIFFUSE ? 4 : 3) bounce_kernel(const PassArgs a, const int bounce) {
    constexpr bool kDiffuse = KIND == Q_DIFFUSE;
    pdl_launch_dependents();
    if (LAST && !kDiffuse) return; // a specular vertex on the last segment contributes nothing
    SceneAccess<false> S;
    S.g = &a.scene;
    pdl_wait();
    const uint32_t n = a.counts[bounce * 4 + KIND];
    if (blockIdx.x * kThreads >= n) return; // short queue: surplus CTAs leave
    const uint32_t* __restrict__ qin = a.q[(bounce & 1) * 3 + (KIND - 1)];
    WarpCursor ray_cur = {0, 0, 0, false};
    uint32_t* const ray_counter = a.counts + bounce * 4 + Q_RAYS;
    unsigned traced = 0, shadow_rays = 0, calls = 0;
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t lane = threadIdx.x & 31u;

    __shared__ VertexStage<!FIRST> stage;
    uint32_t q = blockIdx.x * kThreads + threadIdx.x;
    uint32_t s_cur = kInvalid, s_nxt = kInvalid;
    if (q < n) s_cur = qin[q];
    if (q + stride < n && q + stride >= q) s_nxt = qin[q + stride];
    int buf = 0;
    prefetch_vertex(a, s_cur, stage, 0);
    for (; q - lane < n; q += stride) {
        uint32_t s_nn = kInvalid;
        {
            const uint32_t q2 = q + 2u * stride;
            if (q2 < n && q2 > q) s_nn = qin[q2];
        }
        prefetch_vertex(a, s_nxt, stage, buf ^ 1); // next slot's state, in flight during this body
        cp_async_wait<1>();                        // this slot's state has landed
        const uint32_t slot = s_cur;
        bool ray_shadow = false, ray_cont = false; // rays to enqueue
        float3 ray_o = f3(0.f, 0.f, 0.f), ray_w = f3(0.f, 0.f, 1.f), ray_d = f3(0.f, 0.f, 1.f), ray_rgb = f3(0.f, 0.f, 0.f);
        float ray_tmax = 0.0f;
        uint32_t ray_flags = 0;
        if (slot != kInvalid) {
            ++calls;
            const int tid = threadIdx.x;
            const float4 hp = stage.hp[buf][tid], dw = stage.dw[buf][tid];
            const float3 p = f3(hp.x, hp.y, hp.z), d = f3(dw.x, dw.y, dw.z);
            const uint32_t prim = __float_as_uint(hp.w), pixel = __float_as_uint(dw.w);
            float3 T = f3(1.f, 1.f, 1.f);
            uint32_t sample;
            if (FIRST) {
                uint32_t lp;
                sample = uint32_t(a.sample_base) + fast_div(slot, a.pix_count, lp);
            } else {
                const float4 tp = stage.tp[buf][tid];
                T = f3(tp.x, tp.y, tp.z);
                sample = __float_as_uint(tp.w);
            }
            const Shaded sh = shade_vertex<KIND, LAST, false>(a, S, bounce, p, d, T, prim, pixel, sample);
            T = sh.T;
            const bool cont = !LAST && (T.x > 0.0f || T.y > 0.0f || T.z > 0.0f);
            if (cont) ++traced;
            if (sh.want_shadow) ++shadow_rays;
            ray_o = sh.no; ray_w = sh.w; ray_d = sh.nd; ray_rgb = sh.lit_rgb; ray_tmax = sh.tmax_s;
            ray_flags = (sh.flags & 1u) ? kRaySpecular : 0u;
            // the continuation vertex's direction / pixel / throughput / sample are known now;
            // trace_kernel adds the hit point and primitive (or ends the path)
            ray_shadow = kDiffuse && sh.want_shadow;
            ray_cont = cont;
            if (cont) {
                a.dw[slot] = make_float4(sh.nd.x, sh.nd.y, sh.nd.z, __uint_as_float(pixel));
                a.tp[slot] = make_float4(T.x, T.y, T.z, __uint_as_float(sample));
            }
        }
        // ray records: (origin, tmax) (direction, slot | kRayShadow | kRaySpecular) (light sample rgb)
        const uint32_t is = warp_reserve(ray_cur, ray_shadow, ray_counter);
        if (is != kInvalid) {
            st_stream(a.ray0 + is, ray_o.x, ray_o.y, ray_o.z, __float_as_uint(ray_tmax));
            st_stream(a.ray1 + is, ray_w.x, ray_w.y, ray_w.z, slot | kRayShadow);
            st_stream(a.ray2 + is, ray_rgb.x, ray_rgb.y, ray_rgb.z, 0u);
            if (a.rkey) a.rkey[is] = (uint16_t)ray_sort_key(a.scene, ray_o, ray_w, true);
        }
        const uint32_t ic = warp_reserve(ray_cur, ray_cont, ray_counter);
        if (ic != kInvalid) {
            st_stream(a.ray0 + ic, ray_o.x, ray_o.y, ray_o.z, __float_as_uint(FLT_MAX));
            st_stream(a.ray1 + ic, ray_d.x, ray_d.y, ray_d.z, slot | ray_flags);
            if (a.rkey) a.rkey[ic] = (uint16_t)ray_sort_key(a.scene, ray_o, ray_d, false);
        }
        s_cur = s_nxt; s_nxt = s_nn;
        buf ^= 1;
    }
    cp_async_wait<0>();
    for (uint32_t i = ray_cur.pos + lane; i < ray_cur.end; i += 32u) a.ray1[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInvalid));
    if (ray_cur.has_next) { // a chunk reserved ahead of time and never used
        const uint32_t base = __shfl_sync(kFull, ray_cur.next, 0);
        for (uint32_t i = lane; i < kChunk; i += 32u) a.ray1[base + i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInvalid));
    }
    calls = warp_sum(calls);
    traced = warp_sum(traced);
    shadow_rays = warp_sum(shadow_rays);
    if (lane == 0 && calls) {
        atomicAdd(a.totals + 2, (unsigned long long)calls);
        if (FIRST) atomicAdd(a.totals + 3, (unsigned long long)calls);
        if (traced) atomicAdd(a.totals + 0, (unsigned long long)traced);
        if (shadow_rays) atomicAdd(a.totals + 1, (unsigned long long)shadow_rays);
        if (!FIRST && kDiffuse) atomicAdd(a.totals + 5, (unsigned long long)calls); // diffuse vertices whose light sample may read L
    }
}

// ---- trace (tree scenes): persistent walk with dynamic ray fetch -----------------------------
// One launch per bounce over the ray queue the bounce kernels filled. Every lane owns at most one
// ray; a round = every lane with a ray walks to its next non-empty leaf and tests it (TreeWalk::step).
// Lanes whose ray finished deliver the result -- shadow ray: add the light sample unless occluded;
// continuation ray: hit point + primitive into the vertex record, slot into its material queue --
// and, once PassArgs::refill lanes are idle, the warp pulls that many new rays with ONE atomic. A warp's
// time is then the sum of its rays' rounds / 32, not 32 x the longest walk.

template <bool COOP, int WALK, int OCC = 3> __global__ void __launch_bounds__(kThreads, OCC) trace_kernel(const PassArgs a, const int bounce) {
    pdl_launch_dependents();
    const SceneAccess<false> S = stage_scene<false>(a);
    pdl_wait();
    const uint32_t n = a.counts[bounce * 4 + Q_RAYS];
    if (a.ray_hint && blockIdx.x == 0 && threadIdx.x == 0) atomicMax(a.ray_hint + bounce, n);
    if (n == 0) return;
    const bool sort = bounce + 1 < a.max_depth;
    const bool next_last = bounce + 2 >= a.max_depth;
    Sorter out;
    out.init(a, bounce + 1);
    uint32_t* const fetch = a.counts + (kMaxPathDepth + 1) * 4 + bounce;
    const uint32_t lane = threadIdx.x & 31u;
    typename WalkOf<WALK>::type w = {};
    unsigned cnt_node = 0, cnt_prim = 0;
    bool have = false, more = true;
    uint32_t tag = 0; // slot | flags of the lane's ray
    float3 rgb = f3(0.f, 0.f, 0.f);
    int pend_kind = -1;
    uint32_t pend_slot = 0;
    unsigned lit = 0;
    while (true) {
        // (converged) deliver the slots whose continuation ray found a surface
        if (sort && __any_sync(kFull, pend_kind >= 0)) {
            out.push(pend_kind, pend_slot);
            pend_kind = -1;
        }
        const uint32_t idle = __ballot_sync(kFull, !have);
        if (more && idle != 0u && (__popc(idle) >= a.refill || idle == kFull)) {
            const uint32_t cnt = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(fetch, cnt);
            base = __shfl_sync(kFull, base, 0);
            if (base + cnt >= n) more = false; // the queue is drained (warp-uniform)
            const uint32_t at = base + __popc(idle & ((1u << lane) - 1u));
            if (!have && at < n) {
                const uint32_t i = at < a.perm_n ? __ldg(a.perm + at) : at; // sorted by origin cell (ray_sort.cu) / queue order
                const float4 r1 = ld_stream(a.ray1 + i); // streamed: read once (evict first, the scene stays in L2)
                tag = __float_as_uint(r1.w);
                if (tag != kInvalid) { // not the padding of a producer's last chunk
                    const float4 r0 = ld_stream(a.ray0 + i);
                    const bool shadow = (tag & kRayShadow) != 0u;
                    if (shadow) {
                        const float4 r2 = ld_stream(a.ray2 + i);
                        rgb = f3(r2.x, r2.y, r2.z);
                    }
                    have = true;
                    if (w.init(S, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), r0.w, shadow)) have = false, pend_kind = -2; // finished at once
                }
            }
        }
        // a ray that finished inside init() (missed the root box / one-leaf scene) is delivered like any other
        bool finished = pend_kind == -2;
        if (finished) pend_kind = -1;
        if (__ballot_sync(kFull, have || finished) == 0u) {
            if (!more) break;
            continue;
        }
        if (w.template step<false, COOP>(S, have)) { // all 32 lanes: one round
            finished = true;
            have = false;
        }
        if (finished) {
            if (WALK == 2 || WALK == 6) walk_counts(w, cnt_node, cnt_prim);
            const uint32_t slot = tag & kRaySlotMask;
            if (tag & kRayShadow) {
                if (w.best_prim == kInvalid) { // unoccluded: the light sample counts
                    ++lit;
                    add_radiance(a, slot, rgb);
                }
            } else if (w.best_prim != kInvalid) {
                const float4 t4 = S.hot_row(w.best_prim, 3);
                const int bsdf = __float_as_int(t4.y);
                if (bsdf == G19_BSDF_EMITTER) {
                    // emission counts after specular bounces only (NEE covers the diffuse ones)
                    if (tag & kRaySpecular) {
                        const float4 T = a.tp[slot];
                        const MaterialD& m = a.scene.materials[__float_as_int(t4.x)];
                        add_radiance(a, slot, f3(T.x * m.emission[0], T.y * m.emission[1], T.z * m.emission[2]));
                    }
                } else if (bsdf == G19_BSDF_DIFFUSE || !next_last) {
                    const float3 p = w.o + w.d * w.best;
                    a.hp[slot] = make_float4(p.x, p.y, p.z, __uint_as_float(w.best_prim));
                    pend_kind = bsdf;
                    pend_slot = slot;
                }
            }
        }
    }
    if (sort) out.flush();
    lit = warp_sum(lit);
    if (lane == 0 && lit) atomicAdd(a.totals + 4, (unsigned long long)lit);
    if (WALK == 2 || WALK == 6) flush_walk_counts(a, cnt_node, cnt_prim);
}

// ---- primary-hit AOV ------------------------------------------------------------------------
// The ray the REFERENCE casts -- through the integer pixel corner, no jitter (raytracer.h:41-43) -- traced
// through THIS engine's structures (flat pair loops / linear octree, nearest hit): entity id (push order, the
// numbering REF mode reports), hit point and the normal turned towards the ray (what ImpTriangle::intersect
// returns, entities.h:239-246; spheres: (p - centre) / r, entities.h:94). Ties PATH mode back to the pinned
// mode: the ids must equal REF's wherever "last hit wins" coincides with "nearest hit" (tests/test_path_link.py),
// and with max_depth = 0 these three arrays feed the reference's own shade (ref_shade_kernel), so the default
// scene closes the loop through the PATH pipeline.
template <bool ALL>
__global__ void __launch_bounds__(kThreads, 2) primary_kernel(const PassArgs a, int32_t* __restrict__ ids, double* __restrict__ points,
                                                              double* __restrict__ normals) {
    const SceneAccess<ALL> S = stage_scene<ALL>(a);
    const uint32_t n = uint32_t(a.map.n_local_pix);
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t lp = blockIdx.x * kThreads + threadIdx.x; lp - lane < n; lp += stride) { // warp-uniform trip count
        bool live = false;
        float3 o = f3(0.f, 0.f, 0.f), d = f3(0.f, 0.f, 1.f);
        if (lp < n) {
            const uint32_t lt = lp >> 10, in = lp & 1023u;
            uint32_t tx;
            const uint32_t t = lt * uint32_t(a.map.world) + uint32_t(a.map.rank);
            const uint32_t ty = fast_div(t, uint32_t(a.map.tiles_x), tx);
            const int x = int(tx * kTile + (in & 31u)), y = int(ty * kTile + (in >> 5));
            live = x < a.map.w && y < a.map.h;
            if (live) {
                const PathCamera& c = a.cam;
                const float fx = float(x) * 0.0002f, fy = float(y) * 0.0002f;
                o = f3(c.pos[0], c.pos[1], c.pos[2]);
                d = normalize(f3(c.top_left[0] - c.left[0] * fx - c.up[0] * fy, c.top_left[1] - c.left[1] * fx - c.up[1] * fy,
                                 c.top_left[2] - c.left[2] * fx - c.up[2] * fy));
            }
        }
        float t;
        uint32_t prim;
        const bool hit = nearest<ALL, false, 1>(S, live, o, d, t, prim); // all 32 lanes walk together
        if (lp < n) {
            int32_t id = -1;
            double px = DBL_MAX, py = DBL_MAX, pz = DBL_MAX, nx = 0, ny = 0, nz = 0; // a miss looks like REF's (ref_visibility_kernel)
            if (hit) {
                const float3 p = o + d * t;
                const float4 r0 = S.hot_row(prim, 0), tag = S.hot_row(prim, 3);
                float4 c0, c1;
                S.cold(prim, c0, c1);
                float3 ng = f3(c0.x, c0.y, c0.z);
                int half = 0;
                if (tag.z == 0.0f) {
                    ng = (p - f3(r0.x, r0.y, r0.z)) * __fdividef(1.0f, r0.w);
                } else {
                    if (tag.z > 1.5f) { // merged parallelogram: which of its two triangles (the builder's b1 >= b2 rule)
                        const float4 r1 = S.hot_row(prim, 1);
                        const float b1 = fmaf(r0.x, p.x, fmaf(r0.y, p.y, fmaf(r0.z, p.z, r0.w)));
                        const float b2 = fmaf(r1.x, p.x, fmaf(r1.y, p.y, fmaf(r1.z, p.z, r1.w)));
                        half = b1 >= b2 ? 0 : 1;
                    }
                    if (dot(ng, d) > 0.0f) ng = -ng;
                }
                id = a.scene.prim_entity[2 * size_t(prim) + half];
                px = double(p.x); py = double(p.y); pz = double(p.z);
                nx = double(ng.x); ny = double(ng.y); nz = double(ng.z);
            }
            ids[lp] = id;
            if (points) { points[lp] = px; points[size_t(n) + lp] = py; points[2 * size_t(n) + lp] = pz; }
            if (normals) { normals[lp] = nx; normals[size_t(n) + lp] = ny; normals[2 * size_t(n) + lp] = nz; }
        }
    }
}

// ---- accumulate / resolve ---------------------------------------------------------------
// FLAT: the finished paths' radiance is a float4 per slot, stored once for EVERY slot of the pass by the kernel
// that ended its path, so it is only read here; otherwise three planes that paths add to and this kernel clears.
template <bool FLAT> __global__ void __launch_bounds__(kThreads) accumulate_kernel(const PassArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t npix = (uint32_t)a.map.n_local_pix, nwin = a.pix_count;
    const uint32_t wp = blockIdx.x * kThreads + threadIdx.x; // pixel within the pass's window
    if (wp < nwin) {
        const uint32_t lp = a.pix_base + wp;
        float acc0 = a.accum[lp], acc1 = a.accum[(size_t)npix + lp], acc2 = a.accum[2 * (size_t)npix + lp];
        // strictly in sample order (independent of the pass split); loads batched four samples deep
        if constexpr (FLAT) {
            float4* L = reinterpret_cast<float4*>(a.L) + wp;
            int s = 0;
            for (; s + 4 <= a.spp_pass; s += 4) {
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = L[(size_t)(s + k) * nwin];
#pragma unroll
                for (int k = 0; k < 4; ++k) { acc0 += v[k].x; acc1 += v[k].y; acc2 += v[k].z; }
            }
            for (; s < a.spp_pass; ++s) {
                const float4 v = L[(size_t)s * nwin];
                acc0 += v.x; acc1 += v.y; acc2 += v.z;
            }
        } else {
            float* L0 = a.L + wp;
            float* L1 = a.L + a.plane + wp;
            float* L2 = a.L + 2 * a.plane + wp;
            int s = 0;
            for (; s + 4 <= a.spp_pass; s += 4) {
                float v0[4], v1[4], v2[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    size_t i = (size_t)(s + k) * nwin;
                    v0[k] = L0[i]; v1[k] = L1[i]; v2[k] = L2[i];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    size_t i = (size_t)(s + k) * nwin;
                    acc0 += v0[k]; acc1 += v1[k]; acc2 += v2[k];
                    if (v0[k] != 0.0f) L0[i] = 0.0f;
                    if (v1[k] != 0.0f) L1[i] = 0.0f;
                    if (v2[k] != 0.0f) L2[i] = 0.0f;
                }
            }
            for (; s < a.spp_pass; ++s) {
                size_t i = (size_t)s * nwin;
                float v0 = L0[i], v1 = L1[i], v2 = L2[i];
                acc0 += v0; acc1 += v1; acc2 += v2;
                if (v0 != 0.0f) L0[i] = 0.0f;
                if (v1 != 0.0f) L1[i] = 0.0f;
                if (v2 != 0.0f) L2[i] = 0.0f;
            }
        }
        a.accum[lp] = acc0;
        a.accum[(size_t)npix + lp] = acc1;
        a.accum[2 * (size_t)npix + lp] = acc2;
    }
    // the pass is over: clear its queue lengths for the next one
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < (kMaxPathDepth + 1) * 5; i += kThreads) a.counts[i] = 0; // queue lengths + fetch cursors
}

__global__ void __launch_bounds__(kThreads) resolve_kernel(TileMap map, const float* __restrict__ accum, int spp,
                                                           float* __restrict__ rad_l, uint8_t* __restrict__ rgb_l) {
    const uint32_t npix = (uint32_t)map.n_local_pix;
    uint32_t lp = blockIdx.x * kThreads + threadIdx.x;
    if (lp >= npix) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = __fdiv_rn(accum[(size_t)c * npix + lp], float(spp));
        rad_l[3 * (size_t)lp + c] = v;
        float cl = fminf(fmaxf(v, 0.0f), 1.0f);
        rgb_l[3 * (size_t)lp + c] = (uint8_t)(int)(255.0f * cl); // truncation, as Image::setPixel
    }
}

// Persistent grids: SM count x the CTAs the occupancy calculator says are resident.
template <typename K> static int resident_grid(K kernel, size_t smem, int sm_count) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * sm_count;
}

static bool all_staged(const PassArgs& a) { // one leaf, everything in shared memory
    return a.scene.n_nodes == 1 && a.stage_nodes >= 1 && a.stage_prims >= a.scene.n_index &&
           a.scene.n_index == a.scene.n_prims && a.stage_cold >= a.scene.n_prims && a.stage_lights >= a.scene.n_lights;
}

size_t path_smem_bytes(const PassArgs& a) {
    size_t nb = (size_t(a.stage_nodes) * 8 + 15) & ~size_t(15);
    size_t cb = all_staged(a) ? size_t(a.stage_cold) * 32 + size_t(a.stage_lights) * 80 + size_t(a.scene.pairs_bytes) : 0;
    return nb + size_t(a.stage_prims) * 64 + cb + size_t(a.stack_levels) * kThreads * 4 + 16 +
           (a.coop_leaf ? size_t(kThreads / 32) * 32 * 8 : 0);
}

// per thread: launches are issued and their errors collected on the thread inside path_render, so several
// g19_ctx driven from several threads never see (or race on) each other's failures
static thread_local char g_launch_error[256] = "";
static void note_launch_error(const char* what, cudaError_t e, size_t smem, int grid) {
    if (!g_launch_error[0])
        snprintf(g_launch_error, sizeof g_launch_error, "%s: %s (dynamic smem %zu B, grid %d)", what, cudaGetErrorString(e),
                 smem, grid);
}

// The grid of a persistent kernel = SM count x resident CTAs, cached per (device, kernel, smem size):
// function attributes are per device, and one process may drive several (one g19_ctx each).
template <typename K> static int persistent_grid(K kernel, size_t smem, int sm_count) {
    struct Entry { int device; const void* fn; size_t smem; int grid; };
    static Entry cache[256];
    static int n_cache = 0;
    static std::mutex lock;
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> guard(lock);
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].device == device && cache[i].fn == (const void*)kernel && cache[i].smem == smem) return cache[i].grid;
    // static (prefetch stage) + dynamic (scene prefix, stack) shared memory may exceed 48 KB.
    // The attribute is a CAP: always raise it to the same ceiling, never to this scene's size
    // (a smaller later scene would otherwise lower it under a cached larger launch).
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) note_launch_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", e, smem, 0);
    const int grid = resident_grid(kernel, smem, sm_count);
    if (n_cache < 256) cache[n_cache++] = Entry{device, (const void*)kernel, smem, grid};
    return grid;
}

// Launch with programmatic stream serialization: the kernel may start while its predecessor in
// the stream drains; it synchronises with pdl_wait() itself.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int KIND, bool FIRST, bool LAST, bool ALL>
static void launch_bounce_k(const PassArgs& a, int bounce, size_t smem, int sm_count, cudaStream_t s) {
    void (*kernel)(PassArgs, int);
    if constexpr (ALL) {
        if constexpr (KIND == Q_DIFFUSE) {
            if (a.bounce_occ == 4) kernel = (a.kind_mask & 6u) ? bounce_flat_kernel<KIND, FIRST, LAST, true, 4> : bounce_flat_kernel<KIND, FIRST, LAST, false, 4>;
            else if (!LAST && a.fold_last) kernel = (a.kind_mask & 6u) ? bounce_flat_kernel<KIND, FIRST, LAST, true, 3, true> : bounce_flat_kernel<KIND, FIRST, LAST, false, 3, true>;
            else kernel = (a.kind_mask & 6u) ? bounce_flat_kernel<KIND, FIRST, LAST, true> : bounce_flat_kernel<KIND, FIRST, LAST, false>;
        }
        else if (!LAST && a.fold_last) kernel = bounce_flat_kernel<KIND, FIRST, LAST, true, 3, true>;
        else kernel = bounce_flat_kernel<KIND, FIRST, LAST, true>;
    } else {
        kernel = bounce_kernel<KIND, FIRST, LAST>;
    }
    const int grid = persistent_grid(kernel, smem, sm_count);
    cudaError_t e = launch_pdl(kernel, grid, smem, s, a, bounce);
    if (e != cudaSuccess) note_launch_error("bounce kernel launch", e, smem, grid);
}

template <int KIND, bool ALL> static void launch_bounce_fl(const PassArgs& a, int bounce, size_t smem, int sm_count, cudaStream_t s) {
    const bool first = bounce == 0, last = bounce + 1 >= a.max_depth;
    if (first && last) launch_bounce_k<KIND, true, true, ALL>(a, bounce, smem, sm_count, s);
    else if (first) launch_bounce_k<KIND, true, false, ALL>(a, bounce, smem, sm_count, s);
    else if (last) launch_bounce_k<KIND, false, true, ALL>(a, bounce, smem, sm_count, s);
    else launch_bounce_k<KIND, false, false, ALL>(a, bounce, smem, sm_count, s);
}

} // namespace

void launch_raygen_extend(const PassArgs& a, int sm_count, cudaStream_t s) {
    const size_t smem = path_smem_bytes(a);
    cudaError_t e;
    int grid;
    if (all_staged(a)) {
        auto kernel = (a.kind_mask & 6u) ? raygen_extend_flat_kernel<true> : raygen_extend_flat_kernel<false>;
        grid = persistent_grid(kernel, smem, sm_count);
        e = launch_pdl(kernel, grid, smem, s, a);
    } else {
        const int occ = a.raygen_occ; // tuning knob
        // 3 CTAs per SM at 80 registers (68 bytes of spills) beat 2 at 95: 23.8 vs 26.4 ms per 66 M camera rays
        void (*kernel)(PassArgs);
        if (a.walk == 6) kernel = raygen_extend_kernel<3, false, 6>;
        else if (a.walk == 5) kernel = raygen_extend_kernel<3, false, 5>;
        else if (a.walk == 3) kernel = raygen_extend_kernel<3, false, 3>;
        else if (a.walk == 2) kernel = a.coop_leaf ? raygen_extend_kernel<3, true, 2> : raygen_extend_kernel<3, false, 2>;
        else if (a.walk == 1) kernel = a.coop_leaf ? (occ == 2 ? raygen_extend_kernel<2, true, 1> : raygen_extend_kernel<3, true, 1>)
                                                   : (occ == 2 ? raygen_extend_kernel<2, false, 1> : raygen_extend_kernel<3, false, 1>);
        else kernel = a.coop_leaf ? (occ == 2 ? raygen_extend_kernel<2, true, 0> : raygen_extend_kernel<3, true, 0>)
                                  : (occ == 2 ? raygen_extend_kernel<2, false, 0> : raygen_extend_kernel<3, false, 0>);
        grid = persistent_grid(kernel, smem, sm_count);
        e = launch_pdl(kernel, grid, smem, s, a);
    }
    if (e != cudaSuccess) note_launch_error("raygen_extend kernel launch", e, smem, grid);
}

void launch_trace(const PassArgs& a, int bounce, int sm_count, cudaStream_t s) {
    const size_t smem = path_smem_bytes(a);
    void (*kernel)(PassArgs, int);
    if (a.walk == 6) kernel = trace_kernel<false, 6, 3>;
    else if (a.walk == 5) kernel = a.trace_occ == 4 ? trace_kernel<false, 5, 4> : trace_kernel<false, 5, 3>;
    else if (a.walk == 3) kernel = a.trace_occ == 4 ? trace_kernel<false, 3, 4> : trace_kernel<false, 3, 3>;
    else if (a.walk == 2) kernel = a.coop_leaf ? trace_kernel<true, 2> : trace_kernel<false, 2>;
    else if (a.walk == 1) kernel = a.coop_leaf ? (a.trace_occ == 4 ? trace_kernel<true, 1, 4> : trace_kernel<true, 1, 3>) : trace_kernel<false, 1>;
    else kernel = a.coop_leaf ? trace_kernel<true, 0> : trace_kernel<false, 0>;
    const int grid = persistent_grid(kernel, smem, sm_count);
    cudaError_t e = launch_pdl(kernel, grid, smem, s, a, bounce);
    if (e != cudaSuccess) note_launch_error("trace kernel launch", e, smem, grid);
}

bool path_scene_is_flat(const PassArgs& a) { return all_staged(a); }

// Diffuse-only flat scenes: raygen + first bounce in one launch; false = not applicable (the caller launches the two).
bool launch_bounce_first_fused(const PassArgs& a, int sm_count, cudaStream_t s) {
    if (!all_staged(a) || a.bounce_occ == 4) return false;
    const size_t smem = path_smem_bytes(a);
    void (*kernel)(PassArgs);
    if (a.kind_mask & 6u) // mirror / glass in the scene: diffuse camera hits shaded in place, specular ones queued
        kernel = a.max_depth <= 1 ? bounce_flat_fused_kernel<true, false, true>
                                  : (a.fold_last ? bounce_flat_fused_kernel<false, true, true> : bounce_flat_fused_kernel<false, false, true>);
    else
        kernel = a.max_depth <= 1 ? bounce_flat_fused_kernel<true> : (a.fold_last ? bounce_flat_fused_kernel<false, true> : bounce_flat_fused_kernel<false>);
    const int grid = persistent_grid(kernel, smem, sm_count);
    cudaError_t e = launch_pdl(kernel, grid, smem, s, a);
    if (e != cudaSuccess) note_launch_error("fused first bounce kernel launch", e, smem, grid);
    return true;
}

// Flat scenes with specular materials: one launch serves the three queues of a bounce (not the last).
bool launch_bounce_merged(const PassArgs& a, int bounce, int sm_count, cudaStream_t s) {
    const bool first = bounce == 0, last = bounce + 1 >= a.max_depth;
    if (!all_staged(a) || last || !(a.kind_mask & 6u)) return false;
    const size_t smem = path_smem_bytes(a);
    void (*kernel)(PassArgs, int) = a.fold_last ? (first ? bounce_flat_all_kernel<true, true> : bounce_flat_all_kernel<false, true>)
                                                : (first ? bounce_flat_all_kernel<true> : bounce_flat_all_kernel<false>);
    const int grid = persistent_grid(kernel, smem, sm_count);
    cudaError_t e = launch_pdl(kernel, grid, smem, s, a, bounce);
    if (e != cudaSuccess) note_launch_error("merged bounce kernel launch", e, smem, grid);
    return true;
}

bool launch_bounce(const PassArgs& a, int bounce, int kind, int sm_count, cudaStream_t s) {
    const bool last = bounce + 1 >= a.max_depth;
    if (last && kind != Q_DIFFUSE) return false; // nothing to do: no launch
    const bool all = all_staged(a);
    const size_t smem = all ? path_smem_bytes(a) : 0; // tree scenes only shade here: no scene staging
    switch (kind) {
    case Q_DIFFUSE:
        if (all) launch_bounce_fl<Q_DIFFUSE, true>(a, bounce, smem, sm_count, s);
        else launch_bounce_fl<Q_DIFFUSE, false>(a, bounce, smem, sm_count, s);
        break;
    case Q_MIRROR:
        if (all) launch_bounce_fl<Q_MIRROR, true>(a, bounce, smem, sm_count, s);
        else launch_bounce_fl<Q_MIRROR, false>(a, bounce, smem, sm_count, s);
        break;
    default:
        if (all) launch_bounce_fl<Q_GLASS, true>(a, bounce, smem, sm_count, s);
        else launch_bounce_fl<Q_GLASS, false>(a, bounce, smem, sm_count, s);
        break;
    }
    return true;
}

const char* path_launch_error() { return g_launch_error[0] ? g_launch_error : nullptr; }
void path_clear_launch_error() { g_launch_error[0] = 0; }

void launch_accumulate(const PassArgs& a, cudaStream_t s) {
    int blocks = int((a.pix_count + kThreads - 1) / kThreads);
    if (blocks < 1) blocks = 1;
    cudaError_t e = all_staged(a) ? launch_pdl(accumulate_kernel<true>, blocks, 0, s, a) : launch_pdl(accumulate_kernel<false>, blocks, 0, s, a);
    if (e != cudaSuccess) note_launch_error("accumulate kernel launch", e, 0, blocks);
}

void launch_primary(const PassArgs& a, int32_t* ids_l, double* points_l, double* normals_l, int sm_count, cudaStream_t s) {
    if (a.map.n_local_pix == 0) return;
    PassArgs b = a;
    b.coop_leaf = 0; // sequential leaf tests: no cooperative result slots needed
    b.leaf_batch = 4;
    const size_t smem = path_smem_bytes(b);
    void (*kernel)(PassArgs, int32_t*, double*, double*) = all_staged(b) ? primary_kernel<true> : primary_kernel<false>;
    int grid = persistent_grid(kernel, smem, sm_count);
    grid = std::min(grid, (a.map.n_local_pix + kThreads - 1) / kThreads);
    kernel<<<grid, kThreads, smem, s>>>(b, ids_l, points_l, normals_l);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) note_launch_error("primary kernel launch", e, smem, grid);
}

void launch_resolve(const TileMap& map, const float* accum, int spp, float* rad_l, uint8_t* rgb_l, cudaStream_t s) {
    if (map.n_local_pix == 0) return;
    resolve_kernel<<<(map.n_local_pix + kThreads - 1) / kThreads, kThreads, 0, s>>>(map, accum, spp, rad_l, rgb_l);
}

} // namespace g19
