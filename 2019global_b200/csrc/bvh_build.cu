// bvh_build.cu -- PATH mode, tree scenes: a binary bounding-volume hierarchy over the primitives, built on the GPU.
//
// Why a second structure next to the linear octree (tree_build.cu): the end-state capture of the octree walk
// (profiles/r02g_*) shows 6.9 leaf references per primitive, 17.8 primitive tests and 15 node records per ray, and a
// walk bound by the latency of three dependent loads per leaf test. A BVH references every primitive ONCE (its leaves
// hold 1-4 primitives, stored contiguously in leaf order, so a leaf test is one load chain), and a ray stops descending
// wherever a box lies beyond its nearest hit so far. The reference counterpart is still Octree::intersect
// (reference include/octree.h:132-155): "which entities can this ray hit" -- the answer here is exact nearest-hit, checked
// against the brute-force oracle like the octree's.
//
// Build = LBVH (Lauterbach et al. 2009 / Karras 2012), all on the device, level-free:
//   1. 30-bit Morton code of every primitive's box centre on the root box's 1024^3 grid
//   2. radix sort of (code, primitive id)                                             (CUB)
//   3. one thread per internal node: its key range and split from common-prefix lengths (ties broken by position)
//   4. bottom-up boxes: a leaf thread climbs while it is the second child to arrive      (atomic flags)
//   5. node records for the walk: a child whose range holds <= kBvhLeafMax primitives becomes a LEAF reference
//      (first position, count) -- its own subtree is never visited
//   6. the primitives' 64-byte intersection records gathered into leaf order, original id in the spare word
// Primitives much larger than their neighbours (walls around a mesh) would drag a huge box through every level above
// them; the host keeps those out (path.cu: the "big" list, tested up front by BvhWalk::init).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <string>

#include "path.h"

namespace g19 {

namespace {

constexpr uint32_t kLeafRef = 0x80000000u;

__device__ __forceinline__ uint32_t expand10(uint32_t v) { // bit k -> bit 3k
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void bvh_morton_kernel(const float* __restrict__ boxes, const uint32_t* __restrict__ ids, uint32_t n, float3 lo, float3 inv,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = ids[k];
    const float* b = boxes + 6 * size_t(id);
    const float cx = 0.5f * (b[0] + b[3]), cy = 0.5f * (b[1] + b[4]), cz = 0.5f * (b[2] + b[5]);
    const uint32_t x = uint32_t(min(max(int((cx - lo.x) * inv.x), 0), 1023));
    const uint32_t y = uint32_t(min(max(int((cy - lo.y) * inv.y), 0), 1023));
    const uint32_t z = uint32_t(min(max(int((cz - lo.z) * inv.z), 0), 1023));
    keys[k] = expand10(x) | (expand10(y) << 1) | (expand10(z) << 2);
    vals[k] = id;
}

// common prefix of the keys at positions i and j (64-bit: key, then position, so equal keys still split)
__device__ __forceinline__ int prefix_len(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    if (a != b) return __clz(int(a ^ b));
    return 32 + __clz(i ^ j);
}

// Karras 2012, one thread per internal node i in [0, n - 1). Leaves are referred to as kLeafRef | position.
__global__ void bvh_topology_kernel(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ child, uint32_t* __restrict__ parent_of_node,
                                    uint32_t* __restrict__ parent_of_leaf, int2* __restrict__ range) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = prefix_len(keys, n, i, i + 1) > prefix_len(keys, n, i, i - 1) ? 1 : -1;
    const int dmin = prefix_len(keys, n, i, i - d);
    int lmax = 2;
    while (prefix_len(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (prefix_len(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = prefix_len(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (prefix_len(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const uint32_t left = first == gamma ? (kLeafRef | uint32_t(gamma)) : uint32_t(gamma);
    const uint32_t right = last == gamma + 1 ? (kLeafRef | uint32_t(gamma + 1)) : uint32_t(gamma + 1);
    child[2 * i] = left;
    child[2 * i + 1] = right;
    range[i] = make_int2(first, last);
    if (left & kLeafRef) parent_of_leaf[gamma] = uint32_t(i); else parent_of_node[gamma] = uint32_t(i);
    if (right & kLeafRef) parent_of_leaf[gamma + 1] = uint32_t(i); else parent_of_node[gamma + 1] = uint32_t(i);
    if (i == 0) parent_of_node[0] = 0xffffffffu;
}

struct Box6 {
    float lo[3], hi[3];
};
__device__ __forceinline__ Box6 prim_box(const float* __restrict__ boxes, uint32_t id) {
    const float* b = boxes + 6 * size_t(id);
    Box6 r;
    r.lo[0] = b[0]; r.lo[1] = b[1]; r.lo[2] = b[2];
    r.hi[0] = b[3]; r.hi[1] = b[4]; r.hi[2] = b[5];
    return r;
}

__device__ __forceinline__ Box6 load_box_cg(const Box6* p) {
    const float* f = reinterpret_cast<const float*>(p);
    Box6 r;
    r.lo[0] = __ldcg(f); r.lo[1] = __ldcg(f + 1); r.lo[2] = __ldcg(f + 2);
    r.hi[0] = __ldcg(f + 3); r.hi[1] = __ldcg(f + 4); r.hi[2] = __ldcg(f + 5);
    return r;
}

// bottom-up boxes of the internal nodes: the second child to arrive at a node computes it and climbs on
__global__ void bvh_refit_kernel(const float* __restrict__ boxes, const uint32_t* __restrict__ vals, int n, const uint32_t* __restrict__ child,
                                 const uint32_t* __restrict__ parent_of_node, const uint32_t* __restrict__ parent_of_leaf,
                                 unsigned* __restrict__ arrived, Box6* __restrict__ node_box) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t p = parent_of_leaf[k];
    while (p != 0xffffffffu) {
        __threadfence();
        if (atomicAdd(arrived + p, 1u) == 0u) return; // the sibling subtree is not done yet: its thread takes over
        __threadfence();
        Box6 b;
        for (int c = 0; c < 2; ++c) {
            const uint32_t ref = child[2 * p + c];
            Box6 cb;
            if (ref & kLeafRef) cb = prim_box(boxes, vals[ref & ~kLeafRef]);
            else cb = load_box_cg(node_box + ref); // written by another SM: read past this SM's L1
            if (c == 0) b = cb;
            else
                for (int a = 0; a < 3; ++a) {
                    b.lo[a] = fminf(b.lo[a], cb.lo[a]);
                    b.hi[a] = fmaxf(b.hi[a], cb.hi[a]);
                }
        }
        node_box[p] = b;
        p = parent_of_node[p];
    }
}

// Internal nodes the walk can reach: the root and every node whose range is larger than a leaf. The others (the inside
// of a collapsed leaf) get no record: the records are compacted, in Karras order (which follows the Morton curve), so the
// hierarchy of 1 M primitives is 25 MB of live cache lines instead of 64 MB with dead records in between.
__global__ void bvh_live_kernel(const int2* __restrict__ range, int n, int leaf_max, uint32_t* __restrict__ live) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 rg = range[i];
    live[i] = (i == 0 || rg.y - rg.x + 1 > leaf_max) ? 1u : 0u;
}

// The records the walk reads, 64 B per internal node (Aila & Laine 2009 layout):
//   (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y) (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y) (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z) (ref0, ref1, -, -)
// ref = index of an internal node, or kLeafRef | (count - 1) << 28 | first position in leaf order.
__global__ void bvh_emit_kernel(const float* __restrict__ boxes, const uint32_t* __restrict__ vals, int n, const uint32_t* __restrict__ child,
                                const int2* __restrict__ range, const Box6* __restrict__ node_box, int leaf_max,
                                const uint32_t* __restrict__ live, const uint32_t* __restrict__ new_index, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !live[i]) return;
    Box6 cb[2];
    uint32_t ref[2];
    for (int c = 0; c < 2; ++c) {
        const uint32_t r = child[2 * i + c];
        if (r & kLeafRef) {
            const uint32_t pos = r & ~kLeafRef;
            cb[c] = prim_box(boxes, vals[pos]);
            ref[c] = kLeafRef | pos;
        } else {
            cb[c] = node_box[r];
            const int2 rg = range[r];
            const int cnt = rg.y - rg.x + 1;
            ref[c] = cnt <= leaf_max ? (kLeafRef | (uint32_t(cnt - 1) << 28) | uint32_t(rg.x)) : new_index[r];
        }
    }
    float4* o = out + 4 * size_t(new_index[i]);
    o[0] = make_float4(cb[0].lo[0], cb[0].hi[0], cb[0].lo[1], cb[0].hi[1]);
    o[1] = make_float4(cb[1].lo[0], cb[1].hi[0], cb[1].lo[1], cb[1].hi[1]);
    o[2] = make_float4(cb[0].lo[2], cb[0].hi[2], cb[1].lo[2], cb[1].hi[2]);
    o[3] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), 0.f, 0.f);
}

// ---- 4-wide form: every second level of the binary tree is folded into its parent ---------------------------
// The walk is bound by the latency of one dependent load per visited node (profiles/r02m: 15.7 warps stalled on long
// scoreboard per issue); a node with four children halves the number of those round trips per ray for the same number of
// box tests. Kept nodes = live internal nodes at EVEN depth; a live child at odd depth is replaced by its two children.
__global__ void bvh_kept4_kernel(const uint32_t* __restrict__ parent_of_node, const uint32_t* __restrict__ live, int n, uint32_t* __restrict__ kept) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint32_t k = 0;
    if (live[i]) {
        int depth = 0;
        for (uint32_t p = parent_of_node[i]; p != 0xffffffffu; p = parent_of_node[p]) ++depth;
        k = (depth & 1) ? 0u : 1u;
    }
    kept[i] = k;
}

// 64 B per node -- the capture of the binary walk (profiles/r02m) has the L1 data pipe at 81 % of its peak: every lane
// pulls its own 64-byte record through it, four wavefronts per visit; the same four wavefronts now carry FOUR children.
// Boxes are quantised to 16 bits per coordinate on the root box's grid (40 units / 65535 = 0.6 mm on the room scene, a
// hundredth of a mesh triangle), rounded outwards plus one unit, so a quantised box always contains the exact one:
//   row 0: x  (lo0 | lo1 << 16, lo2 | lo3 << 16, hi0 | hi1 << 16, hi2 | hi3 << 16)     row 1: y     row 2: z
//   row 3: the four references
// Slots are filled from the front (at least two); an empty slot carries the reference 0xffffffff, which the walk checks.
__global__ void bvh_emit4_kernel(const float* __restrict__ boxes, const uint32_t* __restrict__ vals, int n, const uint32_t* __restrict__ child,
                                 const int2* __restrict__ range, const Box6* __restrict__ node_box, int leaf_max,
                                 const uint32_t* __restrict__ kept, const uint32_t* __restrict__ new_index, float3 root_lo, float3 to_grid,
                                 uint4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !kept[i]) return;
    Box6 cb[4];
    uint32_t ref[4];
    int m = 0;
    const float inf = __int_as_float(0x7f800000);
    // a child reference of the binary tree -> one slot (leaf / kept node), or the two slots of an absorbed node's children
    auto slot = [&](uint32_t r, bool absorb) {
        if (r & kLeafRef) {
            const uint32_t pos = r & ~kLeafRef;
            cb[m] = prim_box(boxes, vals[pos]);
            ref[m++] = kLeafRef | pos;
            return true;
        }
        const int2 rg = range[r];
        const int cnt = rg.y - rg.x + 1;
        if (cnt <= leaf_max) {
            cb[m] = node_box[r];
            ref[m++] = kLeafRef | (uint32_t(cnt - 1) << 28) | uint32_t(rg.x);
            return true;
        }
        if (absorb) return false; // a live node at odd depth: the caller takes its children instead
        cb[m] = node_box[r];
        ref[m++] = new_index[r];
        return true;
    };
    for (int c = 0; c < 2; ++c) {
        const uint32_t r = child[2 * i + c];
        if (!slot(r, true)) {
            slot(child[2 * r], false);
            slot(child[2 * r + 1], false);
        }
    }
    for (; m < 4;) {
        for (int a = 0; a < 3; ++a) { cb[m].lo[a] = inf; cb[m].hi[a] = -inf; }
        ref[m++] = 0xffffffffu;
    }
    const float rl[3] = {root_lo.x, root_lo.y, root_lo.z}, tg[3] = {to_grid.x, to_grid.y, to_grid.z};
    uint4 row[3];
    for (int a = 0; a < 3; ++a) {
        uint32_t lo[4], hi[4];
        for (int c = 0; c < 4; ++c) {
            if (ref[c] == 0xffffffffu) { lo[c] = 65535u; hi[c] = 0u; continue; }
            const int ql = int(floorf((cb[c].lo[a] - rl[a]) * tg[a])) - 1, qh = int(ceilf((cb[c].hi[a] - rl[a]) * tg[a])) + 1;
            lo[c] = uint32_t(min(max(ql, 0), 65535));
            hi[c] = uint32_t(min(max(qh, 0), 65535));
        }
        row[a] = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16));
    }
    uint4* o = out + 4 * size_t(new_index[i]);
    o[0] = row[0];
    o[1] = row[1];
    o[2] = row[2];
    o[3] = make_uint4(ref[0], ref[1], ref[2], ref[3]);
}

__global__ void bvh_gather_kernel(const float4* __restrict__ hot, const uint32_t* __restrict__ vals, int n, float4* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = vals[k];
    const float4* src = hot + 4 * size_t(id);
    float4* dst = out + 4 * size_t(k);
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
    float4 tag = src[3];
    tag.w = __uint_as_float(id); // the primitive id the rest of the pipeline knows
    dst[3] = tag;
}

} // namespace

// d_ids: the n primitives that go into the hierarchy (the host kept the big ones out). Returns the root reference
// (a node index, or a leaf reference when n <= leaf_max) in *root; n == 0: *root = 0xffffffff.
int path_build_bvh_device(const float* d_boxes, const uint32_t* d_ids, uint32_t n, const float root_lo[3], const float root_size[3],
                          const PrimHot* d_hot, int leaf_max, cudaStream_t s, DeviceArray& nodes, DeviceArray& nodes4, DeviceArray& prims,
                          uint32_t* root, std::string& err) {
    *root = 0xffffffffu;
    if (n == 0) return G19_OK;
    auto fail = [&](const char* what, cudaError_t e) {
        err = std::string("path_build_bvh_device (") + what + "): " + cudaGetErrorString(e);
        return G19_ERR_CUDA;
    };
    PoolArray keys, vals, keys2, vals2, tmp, child, par_n, par_l, range, arrived, nbox, live, new_index, tmp2; // temporaries: pooled
    auto release_all = [&] {
        for (PoolArray* d : {&keys, &vals, &keys2, &vals2, &tmp, &child, &par_n, &par_l, &range, &arrived, &nbox, &live, &new_index, &tmp2})
            d->release();
    };
    cudaError_t e;
#define BVH_TRY(what, call)            \
    if ((e = (call)) != cudaSuccess) { \
        release_all();                 \
        return fail(what, e);          \
    }
    BVH_TRY("alloc", keys.ensure(size_t(n) * 4, s));
    BVH_TRY("alloc", vals.ensure(size_t(n) * 4, s));
    BVH_TRY("alloc", keys2.ensure(size_t(n) * 4, s));
    BVH_TRY("alloc", vals2.ensure(size_t(n) * 4, s));
    const int threads = 256, blocks = int((n + threads - 1) / threads);
    const float3 lo = make_float3(root_lo[0], root_lo[1], root_lo[2]);
    const float3 inv = make_float3(1024.0f / root_size[0], 1024.0f / root_size[1], 1024.0f / root_size[2]);
    bvh_morton_kernel<<<blocks, threads, 0, s>>>(d_boxes, d_ids, n, lo, inv, static_cast<uint32_t*>(keys.p), static_cast<uint32_t*>(vals.p));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, static_cast<const uint32_t*>(keys.p), static_cast<uint32_t*>(keys2.p),
                                    static_cast<const uint32_t*>(vals.p), static_cast<uint32_t*>(vals2.p), n, 0, 30, s);
    BVH_TRY("alloc", tmp.ensure(tmp_bytes, s));
    BVH_TRY("sort", cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, static_cast<const uint32_t*>(keys.p), static_cast<uint32_t*>(keys2.p),
                                                    static_cast<const uint32_t*>(vals.p), static_cast<uint32_t*>(vals2.p), n, 0, 30, s));
    const uint32_t* skeys = static_cast<const uint32_t*>(keys2.p);
    const uint32_t* svals = static_cast<const uint32_t*>(vals2.p);
    // the records in leaf order
    BVH_TRY("alloc", prims.ensure(size_t(n) * 64));
    bvh_gather_kernel<<<blocks, threads, 0, s>>>(reinterpret_cast<const float4*>(d_hot), svals, int(n), static_cast<float4*>(prims.p));
    if (n <= uint32_t(leaf_max)) { // one leaf
        *root = kLeafRef | ((n - 1) << 28);
        BVH_TRY("nodes", nodes.ensure(64));
        BVH_TRY("nodes", nodes4.ensure(64));
        BVH_TRY("sync", cudaStreamSynchronize(s));
        release_all();
        return G19_OK;
    }
    BVH_TRY("alloc", child.ensure(size_t(n - 1) * 8, s));
    BVH_TRY("alloc", par_n.ensure(size_t(n - 1) * 4, s));
    BVH_TRY("alloc", par_l.ensure(size_t(n) * 4, s));
    BVH_TRY("alloc", range.ensure(size_t(n - 1) * 8, s));
    BVH_TRY("alloc", arrived.ensure(size_t(n - 1) * 4, s));
    BVH_TRY("alloc", nbox.ensure(size_t(n - 1) * sizeof(Box6), s));
    BVH_TRY("alloc", live.ensure(size_t(n) * 4, s));
    BVH_TRY("alloc", new_index.ensure(size_t(n) * 4, s));
    BVH_TRY("memset", cudaMemsetAsync(arrived.p, 0, size_t(n - 1) * 4, s));
    bvh_topology_kernel<<<blocks, threads, 0, s>>>(skeys, int(n), static_cast<uint32_t*>(child.p), static_cast<uint32_t*>(par_n.p),
                                                   static_cast<uint32_t*>(par_l.p), static_cast<int2*>(range.p));
    bvh_refit_kernel<<<blocks, threads, 0, s>>>(d_boxes, svals, int(n), static_cast<const uint32_t*>(child.p), static_cast<const uint32_t*>(par_n.p),
                                                static_cast<const uint32_t*>(par_l.p), static_cast<unsigned*>(arrived.p), static_cast<Box6*>(nbox.p));
    // compaction: position of every live node among the live ones
    BVH_TRY("memset", cudaMemsetAsync(live.p, 0, size_t(n) * 4, s));
    bvh_live_kernel<<<blocks, threads, 0, s>>>(static_cast<const int2*>(range.p), int(n), leaf_max, static_cast<uint32_t*>(live.p));
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<const uint32_t*>(live.p), static_cast<uint32_t*>(new_index.p), int(n), s);
    BVH_TRY("alloc", tmp2.ensure(scan_bytes, s));
    BVH_TRY("scan", cub::DeviceScan::ExclusiveSum(tmp2.p, scan_bytes, static_cast<const uint32_t*>(live.p), static_cast<uint32_t*>(new_index.p), int(n), s));
    uint32_t n_live = 0; // live[n - 1] is 0 (there are n - 1 internal nodes): the last scan entry is the total
    BVH_TRY("copy", cudaMemcpyAsync(&n_live, static_cast<const uint32_t*>(new_index.p) + (n - 1), 4, cudaMemcpyDeviceToHost, s));
    BVH_TRY("sync", cudaStreamSynchronize(s));
    BVH_TRY("alloc", nodes.ensure(size_t(std::max<uint32_t>(n_live, 1u)) * 64));
    bvh_emit_kernel<<<blocks, threads, 0, s>>>(d_boxes, svals, int(n), static_cast<const uint32_t*>(child.p), static_cast<const int2*>(range.p),
                                               static_cast<const Box6*>(nbox.p), leaf_max, static_cast<const uint32_t*>(live.p),
                                               static_cast<const uint32_t*>(new_index.p), static_cast<float4*>(nodes.p));
    // the 4-wide form over the same topology, boxes and leaf order (root = node 0 in both)
    bvh_kept4_kernel<<<blocks, threads, 0, s>>>(static_cast<const uint32_t*>(par_n.p), static_cast<const uint32_t*>(live.p), int(n),
                                                static_cast<uint32_t*>(arrived.p));
    BVH_TRY("memset", cudaMemsetAsync(live.p, 0, size_t(n) * 4, s)); // reuse: live <- kept (entry n - 1 stays 0)
    BVH_TRY("copy", cudaMemcpyAsync(live.p, arrived.p, size_t(n - 1) * 4, cudaMemcpyDeviceToDevice, s));
    BVH_TRY("scan", cub::DeviceScan::ExclusiveSum(tmp2.p, scan_bytes, static_cast<const uint32_t*>(live.p), static_cast<uint32_t*>(new_index.p), int(n), s));
    uint32_t n_kept = 0;
    BVH_TRY("copy", cudaMemcpyAsync(&n_kept, static_cast<const uint32_t*>(new_index.p) + (n - 1), 4, cudaMemcpyDeviceToHost, s));
    BVH_TRY("sync", cudaStreamSynchronize(s));
    BVH_TRY("alloc", nodes4.ensure(size_t(std::max<uint32_t>(n_kept, 1u)) * 64));
    const float3 to_grid = make_float3(65535.0f / root_size[0], 65535.0f / root_size[1], 65535.0f / root_size[2]);
    bvh_emit4_kernel<<<blocks, threads, 0, s>>>(d_boxes, svals, int(n), static_cast<const uint32_t*>(child.p), static_cast<const int2*>(range.p),
                                                static_cast<const Box6*>(nbox.p), leaf_max, static_cast<const uint32_t*>(live.p),
                                                static_cast<const uint32_t*>(new_index.p), lo, to_grid, static_cast<uint4*>(nodes4.p));
    BVH_TRY("launch", cudaGetLastError());
    BVH_TRY("sync", cudaStreamSynchronize(s));
#undef BVH_TRY
    release_all();
    *root = 0u;
    return G19_OK;
}

} // namespace g19
