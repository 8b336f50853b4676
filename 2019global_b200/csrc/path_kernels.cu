// path_kernels.cu -- the wavefront path tracer on sm_100a (G19_MODE_PATH).
//
// One pass traces P = spp_pass * n_local_pix camera paths, bounce by bounce:
//
//   extend<FIRST>   ray generation (bounce 0: regenerated from the pixel index,
//                   no ray record is read) + linear-octree traversal + primitive
//                   intersection; writes the 8-byte hit record; sorts the slot
//                   into a per-material queue by warp-ballot compaction with one
//                   atomic per CTA per queue; emitters are terminated in place
//   shade<KIND>     one launch per material queue (diffuse / mirror / glass):
//                   next-event estimation with an inline any-hit traversal,
//                   BSDF sampling, writes the 40-byte ray record of the next
//                   segment and appends the slot to the next extend queue
//   accumulate      per pixel, in sample order, radiance -> accumulation buffer
//   resolve         mean, clamp, truncating RGB888 store (Image::setPixel,
//                   reference include/image.h:14-16)
//
// State lives in vectorised SoA buffers indexed by path slot (PassArgs);
// queues hold slot indices. Every launch is a persistent grid (SM count x
// resident CTAs) that reads its queue length from device memory, so a frame is
// enqueued without a single host round trip. The RNG is Philox4x32-10 keyed on
// (global pixel, sample, bounce, stream): results do not depend on queue
// order, pass size or how tiles are split across GPUs.
//
// Scene access: the breadth-first prefix of the node array, of the leaf index
// list and of the primitive records is staged into shared memory at kernel
// start with cp.async.bulk (TMA bulk copy, one mbarrier); anything beyond the
// staged prefix is read through L2 with read-only loads.
#include "path.h"

#include <cfloat>

namespace g19 {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kRayEps = 1.0e-3f; // origin offset along the normal (scene units)
constexpr float kPi = 3.14159265358979323846f;

// ---- Philox4x32-10 (Salmon et al. 2011); identical integer stream in oracle/path_oracle.c
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0) {
    uint32_t k1 = 0x32303139u; // "2019"
#pragma unroll
    for (int r = 0; r < 10; ++r) {
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
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize(float3 v) { return v * (1.0f / sqrtf(dot(v, v))); }

// ---- shared-memory scene prefix ----------------------------------------------
// Dynamic shared memory, sized on the host to what the scene needs (PassArgs::stage_*):
//   [nodes: stage_nodes x 8 B][leaf index: stage_index x 4 B][hot prims: stage_prims x 64 B]
//   [traversal stack: stack_levels x kThreads x 4 B][mbarrier: 8 B]
// A Cornell box stages whole (a few hundred bytes) and leaves the SM free for more CTAs.
extern __shared__ __align__(16) unsigned char g19_dyn_smem[];

template <bool ALL> struct SceneAccess {
    const PathSceneD* g;
    const uint2* nodes_s;
    const uint32_t* index_s;
    const float4* hot_s;
    uint32_t* stack;
    int n_nodes_s, n_index_s, n_prims_s;
    __device__ __forceinline__ uint2 node(uint32_t i) const {
        if (ALL || i < (uint32_t)n_nodes_s) return nodes_s[i];
        return __ldg(reinterpret_cast<const uint2*>(g->nodes) + i);
    }
    __device__ __forceinline__ uint32_t prim_index(uint32_t i) const {
        if (ALL || i < (uint32_t)n_index_s) return index_s[i];
        return __ldg(g->prim_index + i);
    }
    __device__ __forceinline__ void prim(uint32_t i, float4& a, float4& b, float4& c, float4& k) const {
        if (ALL || i < (uint32_t)n_prims_s) {
            a = hot_s[4 * i]; b = hot_s[4 * i + 1]; c = hot_s[4 * i + 2]; k = hot_s[4 * i + 3];
        } else {
            const float4* p = reinterpret_cast<const float4*>(g->hot) + 4 * (size_t)i;
            a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); k = __ldg(p + 3);
        }
    }
    __device__ __forceinline__ float4 prim_tag(uint32_t i) const { // row 3: material, bsdf, kind
        if (ALL || i < (uint32_t)n_prims_s) return hot_s[4 * i + 3];
        return __ldg(reinterpret_cast<const float4*>(g->hot) + 4 * (size_t)i + 3);
    }
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage the scene prefix with TMA bulk copies (cp.async.bulk) completing on one mbarrier.
template <bool ALL> __device__ __forceinline__ SceneAccess<ALL> stage_scene(const PassArgs& a) {
    const PathSceneD& g = a.scene;
    SceneAccess<ALL> acc;
    acc.g = &a.scene;
    acc.n_nodes_s = a.stage_nodes;
    acc.n_index_s = a.stage_index;
    acc.n_prims_s = a.stage_prims;
    // byte counts rounded up to 16 (the device arrays are padded, see DeviceArray::ensure)
    const uint32_t nb = (uint32_t(a.stage_nodes) * 8u + 15u) & ~15u;
    const uint32_t ib = (uint32_t(a.stage_index) * 4u + 15u) & ~15u;
    const uint32_t pb = uint32_t(a.stage_prims) * 64u;
    unsigned char* base = g19_dyn_smem;
    acc.nodes_s = reinterpret_cast<const uint2*>(base);
    acc.index_s = reinterpret_cast<const uint32_t*>(base + nb);
    acc.hot_s = reinterpret_cast<const float4*>(base + nb + ib);
    acc.stack = reinterpret_cast<uint32_t*>(base + nb + ib + pb) + threadIdx.x;
    unsigned long long* barp =
        reinterpret_cast<unsigned long long*>(base + nb + ib + pb + uint32_t(a.stack_levels) * kThreads * 4u);
    const uint32_t bar = smem_addr(barp);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb + ib + pb) : "memory");
        if (nb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base)),
                         "l"(g.nodes), "r"(nb), "r"(bar)
                         : "memory");
        if (ib)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base + nb)),
                         "l"(g.prim_index), "r"(ib), "r"(bar)
                         : "memory");
        if (pb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(base + nb + ib)),
                         "l"(g.hot), "r"(pb), "r"(bar)
                         : "memory");
    }
    // everyone waits for phase 0 of the barrier
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar)
                     : "memory");
    }
    return acc;
}

// ---- primitive intersection ---------------------------------------------------
// Returns t in (tmin, tmax) or -1. Triangles: the ray is mapped into the triangle's own
// (b1, b2, h) frame by the precomputed affine rows -- 6 dot products, one fast division,
// no branches. Spheres: unit direction, discriminant from the perpendicular offset.
__device__ __forceinline__ float hit_prim(float4 a, float4 b, float4 c, float4 k, float3 o, float3 d, float tmin,
                                          float tmax) {
    if (k.z != 0.0f) {
        float oz = fmaf(c.x, o.x, fmaf(c.y, o.y, fmaf(c.z, o.z, c.w)));
        float dz = fmaf(c.x, d.x, fmaf(c.y, d.y, c.z * d.z));
        float t = __fdividef(-oz, dz);
        float ox = fmaf(a.x, o.x, fmaf(a.y, o.y, fmaf(a.z, o.z, a.w)));
        float dx = fmaf(a.x, d.x, fmaf(a.y, d.y, a.z * d.z));
        float oy = fmaf(b.x, o.x, fmaf(b.y, o.y, fmaf(b.z, o.z, b.w)));
        float dy = fmaf(b.x, d.x, fmaf(b.y, d.y, b.z * d.z));
        float u = fmaf(t, dx, ox), v = fmaf(t, dy, oy);
        bool ok = (t > tmin) & (t < tmax) & (u >= 0.0f) & (v >= 0.0f) & (u + v <= 1.0f);
        return ok ? t : -1.0f;
    }
    float3 oc = o - f3(a.x, a.y, a.z);
    float bq = dot(oc, d);
    float3 l = oc - d * bq;
    float disc = a.w * a.w - dot(l, l);
    if (disc < 0.0f) return -1.0f;
    float sq = sqrtf(disc);
    float t0 = -bq - sq, t1 = -bq + sq;
    if (t0 > tmin && t0 < tmax) return t0;
    if (t1 > tmin && t1 < tmax) return t1;
    return -1.0f;
}

// Cell edge: the SAME expression as cell_edge() in path.cu (no FMA contraction).
__device__ __forceinline__ float cell_edge(float root_lo, float size_at_level, uint32_t i) {
    return __fadd_rn(root_lo, __fmul_rn(float(i), size_at_level));
}

// Ordered traversal of the linear octree. Children are visited in the order
// i = 0..7 -> octant i ^ a (a = sign mask of the direction): an octant can only
// be occluded by octants that precede it in this order, so the first leaf hit
// bounds everything behind it. Per level the kernel keeps the child base index
// on a short stack in shared memory (one column per thread, conflict free), a
// 4-bit child counter packed in a register, and the integer cell coordinates.
template <bool ANY, bool ALL>
__device__ bool traverse(const SceneAccess<ALL>& S, float3 o, float3 d, float tmin, float tmax, float& t_hit,
                         uint32_t& prim_hit) {
    const PathSceneD& g = *S.g;
    uint32_t* const stack = S.stack;
    float3 dd = d;
    if (fabsf(dd.x) < 1.0e-20f) dd.x = copysignf(1.0e-20f, dd.x);
    if (fabsf(dd.y) < 1.0e-20f) dd.y = copysignf(1.0e-20f, dd.y);
    if (fabsf(dd.z) < 1.0e-20f) dd.z = copysignf(1.0e-20f, dd.z);
    const float3 inv = f3(1.0f / dd.x, 1.0f / dd.y, 1.0f / dd.z);
    const uint32_t a = (dd.x < 0.0f ? 1u : 0u) | (dd.y < 0.0f ? 2u : 0u) | (dd.z < 0.0f ? 4u : 0u);
    float best = tmax;
    uint32_t best_prim = 0xffffffffu;

    auto leaf = [&](uint32_t first, uint32_t n) -> bool {
        for (uint32_t k = 0; k < n; ++k) {
            uint32_t pi = S.prim_index(first + k);
            float4 qa, qb, qc, qk;
            S.prim(pi, qa, qb, qc, qk);
            float t = hit_prim(qa, qb, qc, qk, o, d, tmin, best);
            if (t >= 0.0f) {
                best = t;
                best_prim = pi;
                if (ANY) return true;
            }
        }
        return false;
    };
    auto slab = [&](float lox, float loy, float loz, float hix, float hiy, float hiz, float& tn, float& tf) {
        float x0 = (lox - o.x) * inv.x, x1 = (hix - o.x) * inv.x;
        float y0 = (loy - o.y) * inv.y, y1 = (hiy - o.y) * inv.y;
        float z0 = (loz - o.z) * inv.z, z1 = (hiz - o.z) * inv.z;
        tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
        tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
        tf = tf * 1.0000005f + 1.0e-6f; // conservative: never cull a cell the ray grazes
        tn = tn - fabsf(tn) * 5.0e-7f - 1.0e-6f;
    };

    uint2 root = S.node(0);
    {
        float tn, tf;
        slab(g.root_lo[0], g.root_lo[1], g.root_lo[2], g.root_lo[0] + g.root_size[0], g.root_lo[1] + g.root_size[1],
             g.root_lo[2] + g.root_size[2], tn, tf);
        if (tn > fminf(tf, best) || tf < tmin) return false;
    }
    if (root.y & kLeafBit) {
        leaf(root.x, root.y & ~kLeafBit);
    } else {
        int level = 0;
        uint32_t ix = 0, iy = 0, iz = 0;
        unsigned long long iters = 0;
        stack[0] = root.x;
        while (true) {
            uint32_t i = uint32_t(iters >> (4 * level)) & 0xFu;
            if (i >= 8u) {
                if (level == 0) break;
                --level;
                ix >>= 1; iy >>= 1; iz >>= 1;
                continue;
            }
            iters += 1ull << (4 * level);
            uint32_t c = i ^ a;
            uint2 rec = S.node(stack[level * kThreads] + c);
            if (rec.y == kLeafBit) continue; // empty octant
            uint32_t cx = 2u * ix + (c & 1u), cy = 2u * iy + ((c >> 1) & 1u), cz = 2u * iz + ((c >> 2) & 1u);
            float scale = __int_as_float((127 - (level + 1)) << 23); // 2^-(level+1), exact
            float sx = g.root_size[0] * scale, sy = g.root_size[1] * scale, sz = g.root_size[2] * scale;
            float tn, tf;
            slab(cell_edge(g.root_lo[0], sx, cx), cell_edge(g.root_lo[1], sy, cy), cell_edge(g.root_lo[2], sz, cz),
                 cell_edge(g.root_lo[0], sx, cx + 1), cell_edge(g.root_lo[1], sy, cy + 1),
                 cell_edge(g.root_lo[2], sz, cz + 1), tn, tf);
            if (tn > fminf(tf, best) || tf < tmin) continue;
            if (rec.y & kLeafBit) {
                if (leaf(rec.x, rec.y & ~kLeafBit)) break;
            } else {
                ++level;
                stack[level * kThreads] = rec.x;
                ix = cx; iy = cy; iz = cz;
                iters &= ~(0xFull << (4 * level));
            }
        }
    }
    t_hit = best;
    prim_hit = best_prim;
    return best_prim != 0xffffffffu;
}

// ---- per-slot helpers ------------------------------------------------------------
__device__ __forceinline__ bool slot_pixel(const PassArgs& a, uint32_t slot, int& x, int& y, uint32_t& sample) {
    uint32_t npix = (uint32_t)a.map.n_local_pix;
    uint32_t s = slot / npix, lp = slot - s * npix;
    sample = (uint32_t)a.sample_base + s;
    uint32_t lt = lp >> 10, in = lp & 1023u;
    uint32_t t = lt * (uint32_t)a.map.world + (uint32_t)a.map.rank;
    uint32_t ty = t / (uint32_t)a.map.tiles_x, tx = t - ty * (uint32_t)a.map.tiles_x;
    x = int(tx * kTile + (in & 31u));
    y = int(ty * kTile + (in >> 5));
    return x < a.map.w && y < a.map.h;
}

// The reference pinhole (raytracer.h:26-30,41) with a uniform jitter inside the pixel.
__device__ __forceinline__ void camera_ray(const PassArgs& a, int x, int y, uint32_t sample, float3& o, float3& d) {
    uint32_t pixel = uint32_t(y) * uint32_t(a.map.w) + uint32_t(x);
    uint4 r = philox(pixel, sample, 0u, 0u, a.seed);
    float fx = (float(x) + u01(r.x)) * 0.0002f, fy = (float(y) + u01(r.y)) * 0.0002f;
    const PathCamera& c = a.cam;
    o = f3(c.pos[0], c.pos[1], c.pos[2]);
    float3 v = f3(c.top_left[0] - c.left[0] * fx - c.up[0] * fy, c.top_left[1] - c.left[1] * fx - c.up[1] * fy,
                  c.top_left[2] - c.left[2] * fx - c.up[2] * fy);
    d = normalize(v);
}

// Warp-ballot compaction into NK queues with ONE atomic per CTA per queue.
template <int NK> struct AppendSmem {
    uint32_t wcount[NK][kWarps];
    uint32_t wbase[NK][kWarps];
};

template <int NK>
__device__ __forceinline__ void block_append(AppendSmem<NK>& sm, int kind, uint32_t value, uint32_t* const (&queue)[NK],
                                             uint32_t* const (&counter)[NK]) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        uint32_t m = __ballot_sync(0xffffffffu, kind == k);
        if (lane == 0) sm.wcount[k][warp] = __popc(m);
        if (kind == k) mine = m;
    }
    __syncthreads();
    if (threadIdx.x < NK) {
        const int k = threadIdx.x;
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) total += sm.wcount[k][w];
        uint32_t base = total ? atomicAdd(counter[k], total) : 0u;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            sm.wbase[k][w] = base;
            base += sm.wcount[k][w];
        }
    }
    __syncthreads();
    if (kind >= 0) queue[kind][sm.wbase[kind][warp] + __popc(mine & ((1u << lane) - 1u))] = value;
}

__device__ __forceinline__ void add_radiance(const PassArgs& a, uint32_t slot, float3 v) {
    float* L = a.L;
    L[slot] += v.x;
    L[a.plane + slot] += v.y;
    L[2 * a.plane + slot] += v.z;
}

// ---- extend ------------------------------------------------------------------------
template <bool FIRST, bool ALL>
__global__ void __launch_bounds__(kThreads, 4) extend_kernel(const PassArgs a, const int bounce) {
    __shared__ AppendSmem<3> app_sm;
    const SceneAccess<ALL> S = stage_scene<ALL>(a);

    const uint32_t n = FIRST ? a.n_slots : a.counts[bounce * 4 + Q_EXTEND];
    const uint32_t* __restrict__ qin = a.q[bounce & 1];
    uint32_t* const queues[3] = {a.q[2], a.q[3], a.q[4]};
    uint32_t* const counters[3] = {a.counts + bounce * 4 + Q_DIFFUSE, a.counts + bounce * 4 + Q_MIRROR,
                                   a.counts + bounce * 4 + Q_GLASS};
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t base = blockIdx.x * kThreads; base < n; base += stride) {
        const uint32_t q = base + threadIdx.x;
        int kind = -1;
        uint32_t slot = 0;
        if (q < n) {
            slot = FIRST ? q : qin[q];
            float3 o, d;
            bool live = true;
            if (FIRST) {
                int x, y;
                uint32_t sample;
                live = slot_pixel(a, slot, x, y, sample);
                if (live) camera_ray(a, x, y, sample, o, d);
            } else {
                float4 r0 = a.ro[slot];
                float2 r1 = a.rd[slot];
                o = f3(r0.x, r0.y, r0.z);
                d = f3(r0.w, r1.x, r1.y);
            }
            if (live) {
                float t = FLT_MAX;
                uint32_t prim = 0xffffffffu;
                bool hit = traverse<false, ALL>(S, o, d, 0.0f, FLT_MAX, t, prim);
                a.hit[slot] = make_uint2(__float_as_uint(t), prim);
                if (hit) {
                    const float4 tag = S.prim_tag(prim); // material class rides in the hot record
                    int bsdf = __float_as_int(tag.y);
                    if (bsdf == G19_BSDF_EMITTER) {
                        // emission counts on camera rays and after specular bounces only (NEE covers the rest)
                        float4 T = FIRST ? make_float4(1.f, 1.f, 1.f, 0.f) : a.tp[slot];
                        if (FIRST || (__float_as_uint(T.w) & 1u)) {
                            const MaterialD& m = a.scene.materials[__float_as_int(tag.x)];
                            add_radiance(a, slot, f3(T.x * m.emission[0], T.y * m.emission[1], T.z * m.emission[2]));
                        }
                    } else {
                        kind = bsdf;
                    }
                }
            }
        }
        block_append<3>(app_sm, kind, slot, queues, counters);
    }
}

// ---- shade ---------------------------------------------------------------------------
__device__ __forceinline__ void onb(float3 n, float3& t, float3& b) { // Duff et al. 2017
    float s = copysignf(1.0f, n.z);
    float a = -1.0f / (s + n.z);
    float bb = n.x * n.y * a;
    t = f3(1.0f + s * n.x * n.x * a, s * bb, -s * n.x);
    b = f3(bb, s + n.y * n.y * a, -n.y);
}

template <int KIND, bool FIRST, bool ALL>
__global__ void __launch_bounds__(kThreads, KIND == Q_DIFFUSE ? 3 : 4) shade_kernel(const PassArgs a, const int bounce) {
    __shared__ AppendSmem<1> app_sm;
    SceneAccess<ALL> S;
    if (KIND == Q_DIFFUSE) S = stage_scene<ALL>(a); // shadow rays traverse

    const uint32_t n = a.counts[bounce * 4 + KIND];
    const uint32_t* __restrict__ qin = a.q[1 + KIND];
    uint32_t* const queues[1] = {a.q[(bounce + 1) & 1]};
    uint32_t* const counters[1] = {a.counts + (bounce + 1) * 4 + Q_EXTEND};
    const bool more = bounce + 1 < a.max_depth;
    unsigned shadow_rays = 0, lit = 0;
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t base = blockIdx.x * kThreads; base < n; base += stride) {
        const uint32_t q = base + threadIdx.x;
        int kind = -1;
        uint32_t slot = 0;
        if (q < n) {
            slot = qin[q];
            int x, y;
            uint32_t sample;
            slot_pixel(a, slot, x, y, sample);
            float3 o, d, T;
            if (FIRST) {
                camera_ray(a, x, y, sample, o, d);
                T = f3(1.f, 1.f, 1.f);
            } else {
                float4 r0 = a.ro[slot];
                float2 r1 = a.rd[slot];
                float4 tp = a.tp[slot];
                o = f3(r0.x, r0.y, r0.z);
                d = f3(r0.w, r1.x, r1.y);
                T = f3(tp.x, tp.y, tp.z);
            }
            const uint2 h = a.hit[slot];
            const float t = __uint_as_float(h.x);
            const uint32_t prim = h.y;
            const float3 p = o + d * t;
            const PrimCold cold = a.scene.cold[prim];
            const MaterialD mat = a.scene.materials[cold.material];
            float3 ng;
            {
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.scene.hot) + 4 * (size_t)prim);
                const float4 q3 = __ldg(reinterpret_cast<const float4*>(a.scene.hot) + 4 * (size_t)prim + 3);
                if (q3.z != 0.0f) ng = f3(cold.n[0], cold.n[1], cold.n[2]);
                else ng = (p - f3(q0.x, q0.y, q0.z)) * (1.0f / q0.w);
            }
            const bool entering = dot(ng, d) < 0.0f;
            const float3 nf = entering ? ng : -ng; // normal on the side the ray arrives from
            const float3 albedo = f3(mat.albedo[0], mat.albedo[1], mat.albedo[2]);
            const uint32_t pixel = uint32_t(y) * uint32_t(a.map.w) + uint32_t(x);
            const uint4 r = philox(pixel, sample, uint32_t(bounce), 1u, a.seed);
            float3 no, nd;
            uint32_t flags = 0;
            if (KIND == Q_DIFFUSE) {
                // next-event estimation: one light, one uniformly sampled point, one any-hit ray
                if (a.scene.n_lights > 0) {
                    float pick = u01(r.x) * float(a.scene.n_lights);
                    int li = min(int(pick), a.scene.n_lights - 1);
                    float u1 = pick - float(li), u2 = u01(r.y);
                    const LightD& lt = a.scene.lights[li];
                    float su = sqrtf(u1);
                    float b1 = su * (1.0f - u2), b2 = su * u2;
                    float3 yl = f3(lt.v0[0] + lt.e1[0] * b1 + lt.e2[0] * b2, lt.v0[1] + lt.e1[1] * b1 + lt.e2[1] * b2,
                                   lt.v0[2] + lt.e1[2] * b1 + lt.e2[2] * b2);
                    float3 w = yl - p;
                    float dist2 = dot(w, w);
                    float dist = sqrtf(dist2);
                    w = w * (1.0f / dist);
                    float cs = dot(nf, w);
                    float cl = fabsf(dot(f3(lt.n[0], lt.n[1], lt.n[2]), w));
                    if (cs > 0.0f && cl > 0.0f && dist > 2.0f * kRayEps) {
                        float tt;
                        uint32_t pp;
                        ++shadow_rays;
                        bool blocked = traverse<true, ALL>(S, p + nf * kRayEps, w, 0.0f, dist - 2.0f * kRayEps, tt, pp);
                        if (!blocked) {
                            ++lit;
                            float gterm = cs * cl * lt.area / (dist2 * lt.pdf_pick) * (1.0f / kPi);
                            add_radiance(a, slot, f3(T.x * albedo.x * lt.emission[0] * gterm,
                                                     T.y * albedo.y * lt.emission[1] * gterm,
                                                     T.z * albedo.z * lt.emission[2] * gterm));
                        }
                    }
                }
                // cosine-weighted bounce: pdf cancels cos/pi, throughput *= albedo
                float u3 = u01(r.z), u4 = u01(r.w);
                float rr = sqrtf(u3), phi = 2.0f * kPi * u4;
                float sp, cp;
                sincosf(phi, &sp, &cp);
                float3 tx, ty;
                onb(nf, tx, ty);
                nd = normalize(tx * (rr * cp) + ty * (rr * sp) + nf * sqrtf(fmaxf(0.0f, 1.0f - u3)));
                no = p + nf * kRayEps;
                T = T * albedo;
            } else if (KIND == Q_MIRROR) {
                nd = normalize(d - nf * (2.0f * dot(d, nf)));
                no = p + nf * kRayEps;
                T = T * albedo;
                flags = 1u;
            } else { // dielectric
                float etai = entering ? 1.0f : mat.ior, etat = entering ? mat.ior : 1.0f;
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
                    nd = normalize(d + nf * (2.0f * cosi));
                    no = p + nf * kRayEps;
                } else {
                    nd = normalize(d * eta + nf * (eta * cosi - cost));
                    no = p - nf * kRayEps;
                }
                T = T * albedo;
                flags = 1u;
            }
            if (more && (T.x > 0.0f || T.y > 0.0f || T.z > 0.0f)) {
                a.ro[slot] = make_float4(no.x, no.y, no.z, nd.x);
                a.rd[slot] = make_float2(nd.y, nd.z);
                a.tp[slot] = make_float4(T.x, T.y, T.z, __uint_as_float(flags));
                kind = 0;
            }
        }
        block_append<1>(app_sm, kind, slot, queues, counters);
    }
    if (KIND == Q_DIFFUSE) {
        for (int off = 16; off > 0; off >>= 1) {
            shadow_rays += __shfl_xor_sync(0xffffffffu, shadow_rays, off);
            lit += __shfl_xor_sync(0xffffffffu, lit, off);
        }
        if ((threadIdx.x & 31) == 0 && shadow_rays) {
            atomicAdd(a.totals + 1, (unsigned long long)shadow_rays);
            atomicAdd(a.totals + 4, (unsigned long long)lit);
        }
    }
}

// ---- accumulate / resolve ---------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) accumulate_kernel(const PassArgs a) {
    const uint32_t npix = (uint32_t)a.map.n_local_pix;
    const uint32_t lp = blockIdx.x * kThreads + threadIdx.x;
    if (lp < npix) {
        float acc0 = a.accum[lp], acc1 = a.accum[(size_t)npix + lp], acc2 = a.accum[2 * (size_t)npix + lp];
        float* L0 = a.L + lp;
        float* L1 = a.L + a.plane + lp;
        float* L2 = a.L + 2 * a.plane + lp;
        // strictly in sample order (independent of the pass split); loads batched four samples deep
        int s = 0;
        for (; s + 4 <= a.spp_pass; s += 4) {
            float v0[4], v1[4], v2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                size_t i = (size_t)(s + k) * npix;
                v0[k] = L0[i]; v1[k] = L1[i]; v2[k] = L2[i];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                size_t i = (size_t)(s + k) * npix;
                acc0 += v0[k]; acc1 += v1[k]; acc2 += v2[k];
                if (v0[k] != 0.0f) L0[i] = 0.0f;
                if (v1[k] != 0.0f) L1[i] = 0.0f;
                if (v2[k] != 0.0f) L2[i] = 0.0f;
            }
        }
        for (; s < a.spp_pass; ++s) {
            size_t i = (size_t)s * npix;
            float v0 = L0[i], v1 = L1[i], v2 = L2[i];
            acc0 += v0; acc1 += v1; acc2 += v2;
            if (v0 != 0.0f) L0[i] = 0.0f;
            if (v1 != 0.0f) L1[i] = 0.0f;
            if (v2 != 0.0f) L2[i] = 0.0f;
        }
        a.accum[lp] = acc0;
        a.accum[(size_t)npix + lp] = acc1;
        a.accum[2 * (size_t)npix + lp] = acc2;
    }
    // fold this pass's queue lengths into the running totals and clear them
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < (kMaxPathDepth + 1) * 4; i += kThreads) {
            uint32_t v = a.counts[i];
            if (v) {
                if ((i & 3) == Q_EXTEND) atomicAdd(a.totals + 0, (unsigned long long)v);
                else {
                    atomicAdd(a.totals + 2, (unsigned long long)v);
                    if (i < 4) atomicAdd(a.totals + 3, (unsigned long long)v);
                }
                a.counts[i] = 0;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) resolve_kernel(TileMap map, const float* __restrict__ accum, int spp,
                                                           float* __restrict__ rad_l, uint8_t* __restrict__ rgb_l) {
    const uint32_t npix = (uint32_t)map.n_local_pix;
    uint32_t lp = blockIdx.x * kThreads + threadIdx.x;
    if (lp >= npix) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = accum[(size_t)c * npix + lp] / float(spp);
        rad_l[3 * (size_t)lp + c] = v;
        float cl = fminf(fmaxf(v, 0.0f), 1.0f);
        rgb_l[3 * (size_t)lp + c] = (uint8_t)(int)(255.0f * cl); // truncation, as Image::setPixel
    }
}

} // namespace

// Persistent grids: SM count x the CTAs the occupancy calculator says are resident.
template <typename K> static int resident_grid(K kernel, size_t smem, int sm_count) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm * sm_count;
}

size_t path_smem_bytes(const PassArgs& a, bool with_scene) {
    if (!with_scene) return 0;
    size_t nb = (size_t(a.stage_nodes) * 8 + 15) & ~size_t(15);
    size_t ib = (size_t(a.stage_index) * 4 + 15) & ~size_t(15);
    return nb + ib + size_t(a.stage_prims) * 64 + size_t(a.stack_levels) * kThreads * 4 + 16;
}

static bool all_staged(const PassArgs& a) {
    return a.stage_nodes >= a.scene.n_nodes && a.stage_index >= a.scene.n_index && a.stage_prims >= a.scene.n_prims;
}

// The grid of a persistent kernel = SM count x resident CTAs, cached per (kernel, smem size).
template <typename K> static void launch_persistent(K kernel, const PassArgs& a, int bounce, size_t smem, int sm_count,
                                                    cudaStream_t s) {
    struct Entry { const void* fn; size_t smem; int grid; };
    static Entry cache[64];
    static int n_cache = 0;
    int grid = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].fn == (const void*)kernel && cache[i].smem == smem) grid = cache[i].grid;
    if (!grid) {
        grid = resident_grid(kernel, smem, sm_count);
        if (n_cache < 64) cache[n_cache++] = Entry{(const void*)kernel, smem, grid};
    }
    kernel<<<grid, kThreads, smem, s>>>(a, bounce);
}

void launch_extend(const PassArgs& a, int bounce, int sm_count, cudaStream_t s) {
    const size_t smem = path_smem_bytes(a, true);
    const bool all = all_staged(a);
    if (bounce == 0) {
        if (all) launch_persistent(extend_kernel<true, true>, a, bounce, smem, sm_count, s);
        else launch_persistent(extend_kernel<true, false>, a, bounce, smem, sm_count, s);
    } else {
        if (all) launch_persistent(extend_kernel<false, true>, a, bounce, smem, sm_count, s);
        else launch_persistent(extend_kernel<false, false>, a, bounce, smem, sm_count, s);
    }
}

template <int KIND> static void launch_shade_k(const PassArgs& a, int bounce, int sm_count, cudaStream_t s) {
    const size_t smem = path_smem_bytes(a, KIND == Q_DIFFUSE);
    const bool all = all_staged(a), first = bounce == 0;
    if (first) {
        if (all) launch_persistent(shade_kernel<KIND, true, true>, a, bounce, smem, sm_count, s);
        else launch_persistent(shade_kernel<KIND, true, false>, a, bounce, smem, sm_count, s);
    } else {
        if (all) launch_persistent(shade_kernel<KIND, false, true>, a, bounce, smem, sm_count, s);
        else launch_persistent(shade_kernel<KIND, false, false>, a, bounce, smem, sm_count, s);
    }
}

void launch_shade(const PassArgs& a, int bounce, int kind, int sm_count, cudaStream_t s) {
    switch (kind) {
    case Q_DIFFUSE: launch_shade_k<Q_DIFFUSE>(a, bounce, sm_count, s); break;
    case Q_MIRROR: launch_shade_k<Q_MIRROR>(a, bounce, sm_count, s); break;
    default: launch_shade_k<Q_GLASS>(a, bounce, sm_count, s); break;
    }
}

void launch_accumulate(const PassArgs& a, cudaStream_t s) {
    int blocks = (a.map.n_local_pix + kThreads - 1) / kThreads;
    if (blocks < 1) blocks = 1;
    accumulate_kernel<<<blocks, kThreads, 0, s>>>(a);
}

void launch_resolve(const TileMap& map, const float* accum, int spp, float* rad_l, uint8_t* rgb_l, cudaStream_t s) {
    if (map.n_local_pix == 0) return;
    resolve_kernel<<<(map.n_local_pix + kThreads - 1) / kThreads, kThreads, 0, s>>>(map, accum, spp, rad_l, rgb_l);
}

} // namespace g19
