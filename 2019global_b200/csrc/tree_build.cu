// tree_build.cu -- the PATH-mode linear octree, built on the GPU.
//
// Reference counterpart: Octree::push_back / Node::push_obj / Node::partition (reference
// include/octree.h:20-43,75-129) -- a pointer tree grown one entity at a time on one CPU thread
// (2.3 s and 201 MB at 1 M triangles), which also loses entities when a node
// splits. REF mode keeps that tree, bug for bug (scene.cpp). PATH mode traverses this engine's
// own tree (path.cu: implicit boxes, every primitive reachable, breadth-first 8-byte records),
// and this file builds exactly that tree level by level on the device:
//
//   per level   refs[R] = (frontier node, primitive) pairs, grouped by node, primitive order kept
//     classify  one thread per ref: 8-bit mask of the child cells its box overlaps (the builder's
//               own cell_edge expression, no FMA contraction) + per-child counters
//     decide    one thread per frontier node: leaf if small / deep / splitting does not pay
//               (children would hold >= 3x the references) -- the host builder's rules
//     scan      exclusive scans give leaves their slice of the index array, interior nodes their
//               block of 8 child records, child cells their slice of the next level's refs, and
//               every ref its rank among its node's refs per child (one scan of 8-lane counters)
//     scatter   refs of leaves go to the index array, refs of interior nodes to their children
//
// The result is BIT-IDENTICAL to the host builder's (same node numbering, same leaf lists):
// tests/test_tree_build.py compares both. One host read-back of three counters per level.
#include "path.h"

#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace g19 {
namespace {

constexpr int kThreads = 256;

struct Lane8 {
    uint32_t v[8];
};
struct Lane8Add {
    __host__ __device__ Lane8 operator()(const Lane8& a, const Lane8& b) const {
        Lane8 r;
        for (int k = 0; k < 8; ++k) r.v[k] = a.v[k] + b.v[k];
        return r;
    }
};
struct MaskToLanes {
    __host__ __device__ Lane8 operator()(uint8_t m) const {
        Lane8 r;
        for (int k = 0; k < 8; ++k) r.v[k] = (m >> k) & 1u;
        return r;
    }
};

struct Frontier { // one entry per node of the level being split
    uint32_t node, ix, iy, iz;
};

struct LevelArgs {
    const float* boxes; // [n_prims][6]: lo.xyz, hi.xyz
    float root_lo[3], root_size[3];
    int level, leaf_max, max_depth;
    uint32_t F, R;
};

// the ONE cell-edge expression (path.cu cell_edge): root_lo + float(i) * root_size * 2^-level
__device__ __forceinline__ float edge(float lo, float size, int level, uint32_t i) {
    return __fadd_rn(lo, __fmul_rn(float(i), ldexpf(size, -level)));
}

__global__ void classify_kernel(LevelArgs a, const Frontier* __restrict__ fr, const uint32_t* __restrict__ f_off,
                                const uint32_t* __restrict__ refs, const uint32_t* __restrict__ ref_f,
                                uint8_t* __restrict__ mask, uint32_t* __restrict__ child_cnt) {
    const uint32_t r = blockIdx.x * kThreads + threadIdx.x;
    if (r >= a.R) return;
    const uint32_t f = ref_f[r];
    const uint32_t n = f_off[f + 1] - f_off[f];
    uint8_t m = 0;
    if (!(n <= (uint32_t)a.leaf_max || a.level >= a.max_depth)) {
        const Frontier nd = fr[f];
        const float* b = a.boxes + 6 * (size_t)refs[r];
        const float lo[3] = {b[0], b[1], b[2]}, hi[3] = {b[3], b[4], b[5]};
        const uint32_t ci[3] = {2u * nd.ix, 2u * nd.iy, 2u * nd.iz};
        // per axis: does the box overlap the low half [e0,e1] / the high half [e1,e2] of the cell
        bool in_lo[3], in_hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float e0 = edge(a.root_lo[k], a.root_size[k], a.level + 1, ci[k]);
            const float e1 = edge(a.root_lo[k], a.root_size[k], a.level + 1, ci[k] + 1u);
            const float e2 = edge(a.root_lo[k], a.root_size[k], a.level + 1, ci[k] + 2u);
            in_lo[k] = !(hi[k] < e0 || lo[k] > e1);
            in_hi[k] = !(hi[k] < e1 || lo[k] > e2);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const bool ox = (c & 1) ? in_hi[0] : in_lo[0], oy = (c & 2) ? in_hi[1] : in_lo[1], oz = (c & 4) ? in_hi[2] : in_lo[2];
            if (ox && oy && oz) {
                m |= uint8_t(1u << c);
                atomicAdd(child_cnt + 8 * (size_t)f + c, 1u);
            }
        }
    }
    mask[r] = m;
}

__global__ void decide_kernel(LevelArgs a, const uint32_t* __restrict__ f_off, uint32_t* __restrict__ child_cnt,
                              uint32_t* __restrict__ leaf_len, uint32_t* __restrict__ interior) {
    const uint32_t f = blockIdx.x * kThreads + threadIdx.x;
    if (f >= a.F) return;
    const uint32_t n = f_off[f + 1] - f_off[f];
    bool leaf = n <= (uint32_t)a.leaf_max || a.level >= a.max_depth;
    if (!leaf) {
        unsigned long long refs = 0;
        for (int c = 0; c < 8; ++c) refs += child_cnt[8 * (size_t)f + c];
        if (refs >= 3ull * n) leaf = true; // splitting must pay
    }
    if (leaf)
        for (int c = 0; c < 8; ++c) child_cnt[8 * (size_t)f + c] = 0;
    leaf_len[f] = leaf ? n : 0u;
    interior[f] = leaf ? 0u : 1u;
}

// node records of this level + the next level's frontier
__global__ void emit_kernel(LevelArgs a, const Frontier* __restrict__ fr, const uint32_t* __restrict__ f_off,
                            const uint32_t* __restrict__ interior, const uint32_t* __restrict__ int_rank,
                            const uint32_t* __restrict__ leaf_off, const uint32_t* __restrict__ child_cnt,
                            uint32_t nodes_size, uint32_t index_size, PathNodeD* __restrict__ nodes,
                            Frontier* __restrict__ next_fr, uint32_t* __restrict__ next_cnt) {
    const uint32_t f = blockIdx.x * kThreads + threadIdx.x;
    if (f >= a.F) return;
    const Frontier nd = fr[f];
    if (!interior[f]) {
        nodes[nd.node] = PathNodeD{index_size + leaf_off[f], (f_off[f + 1] - f_off[f]) | kLeafBit};
        return;
    }
    const uint32_t k = int_rank[f], base = nodes_size + 8u * k;
    nodes[nd.node] = PathNodeD{base, 0u};
    for (uint32_t c = 0; c < 8; ++c) {
        next_fr[8 * (size_t)k + c] = Frontier{base + c, 2u * nd.ix + (c & 1u), 2u * nd.iy + ((c >> 1) & 1u), 2u * nd.iz + ((c >> 2) & 1u)};
        next_cnt[8 * (size_t)k + c] = child_cnt[8 * (size_t)f + c];
    }
}

__global__ void scatter_kernel(LevelArgs a, const uint32_t* __restrict__ f_off, const uint32_t* __restrict__ refs,
                               const uint32_t* __restrict__ ref_f, const uint8_t* __restrict__ mask,
                               const Lane8* __restrict__ pos, const uint32_t* __restrict__ interior,
                               const uint32_t* __restrict__ int_rank, const uint32_t* __restrict__ leaf_off,
                               const uint32_t* __restrict__ next_off, uint32_t index_size, uint32_t* __restrict__ index,
                               uint32_t* __restrict__ next_refs, uint32_t* __restrict__ next_ref_f) {
    const uint32_t r = blockIdx.x * kThreads + threadIdx.x;
    if (r >= a.R) return;
    const uint32_t f = ref_f[r], p = refs[r], first = f_off[f];
    if (!interior[f]) {
        index[index_size + leaf_off[f] + (r - first)] = p;
        return;
    }
    const uint8_t m = mask[r];
    const uint32_t k = int_rank[f];
    const Lane8 here = pos[r], start = pos[first];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (m & (1u << c)) {
            const uint32_t dst = next_off[8 * (size_t)k + c] + (here.v[c] - start.v[c]);
            next_refs[dst] = p;
            next_ref_f[dst] = 8u * k + c;
        }
}

__global__ void iota_kernel(uint32_t n, uint32_t* refs, uint32_t* ref_f) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i < n) {
        refs[i] = i;
        ref_f[i] = 0;
    }
}

inline int blocks(size_t n) { return int((n + kThreads - 1) / kThreads); }

// Growable device array that keeps its contents. Temporaries of the builders come from the device's stream-ordered
// memory pool (cudaMallocAsync): the level loop grows a dozen arrays per level, and with cudaMalloc / cudaFree every one
// of those calls synchronised the device and could stall for 0.1-0.25 s behind earlier frees (a 1 M-triangle upload
// varied between 0.17 s and 2.2 s for this loop alone). The pool keeps what it has handed out before (release threshold
// raised once per device), so a second upload allocates nothing.
struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    cudaError_t reserve(size_t need, size_t keep, cudaStream_t s) {
        if (need <= bytes) return cudaSuccess;
        stream = s;
        size_t cap = bytes ? bytes : 4096;
        while (cap < need) cap *= 2;
        void* q = nullptr;
        cudaError_t e = cudaMallocAsync(&q, cap + 256, s);
        if (e != cudaSuccess) return e;
        if (p && keep) e = cudaMemcpyAsync(q, p, keep, cudaMemcpyDeviceToDevice, s);
        if (p) cudaFreeAsync(p, s); // stream-ordered: after the copy
        p = q;
        bytes = cap;
        return e;
    }
    void release() {
        if (p) cudaFreeAsync(p, stream);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

} // namespace

// Builds nodes[] and index[] on the device from per-primitive boxes (device, [n][6] floats).
// On success *d_nodes / *d_index are cudaMalloc'ed arrays the caller owns.
int path_build_tree_device(const float* d_boxes, uint32_t n_prims, const float root_lo[3], const float root_size[3],
                           int leaf_max, int max_depth, cudaStream_t s, PathNodeD** d_nodes, uint32_t* n_nodes,
                           uint32_t** d_index, uint32_t* n_index, int* tree_depth, std::string& err) {
#define TB_CUDA(call)                                                               \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            err = std::string("tree build: " #call ": ") + cudaGetErrorString(e__); \
            for (Buf* b__ : all) b__->release();                                    \
            return G19_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)
    Buf nodes, index, fr[2], f_off[2], refs[2], ref_f[2], mask, pos, child_cnt, leaf_len, interior, leaf_off, int_rank, next_cnt,
        temp, counters;
    Buf* all[] = {&nodes, &index, &fr[0], &fr[1], &f_off[0], &f_off[1], &refs[0], &refs[1], &ref_f[0], &ref_f[1], &mask,
                  &pos, &child_cnt, &leaf_len, &interior, &leaf_off, &int_rank, &next_cnt, &temp, &counters};
    LevelArgs a;
    a.boxes = d_boxes;
    for (int k = 0; k < 3; ++k) {
        a.root_lo[k] = root_lo[k];
        a.root_size[k] = root_size[k];
    }
    a.leaf_max = leaf_max;
    a.max_depth = max_depth;
    a.level = 0;
    a.F = 1;
    a.R = n_prims;
    uint32_t nodes_size = 1, index_size = 0;
    int depth = 0, cur = 0;
    TB_CUDA(nodes.reserve(sizeof(PathNodeD) * 1024, 0, s));
    TB_CUDA(index.reserve(sizeof(uint32_t) * (size_t(n_prims) + 1024), 0, s));
    TB_CUDA(fr[0].reserve(sizeof(Frontier), 0, s));
    TB_CUDA(f_off[0].reserve(2 * sizeof(uint32_t), 0, s));
    TB_CUDA(refs[0].reserve(sizeof(uint32_t) * size_t(n_prims ? n_prims : 1), 0, s));
    TB_CUDA(ref_f[0].reserve(sizeof(uint32_t) * size_t(n_prims ? n_prims : 1), 0, s));
    TB_CUDA(counters.reserve(4 * sizeof(uint32_t), 0, s));
    {
        const Frontier root = {0, 0, 0, 0};
        const uint32_t off[2] = {0, n_prims};
        TB_CUDA(cudaMemcpyAsync(fr[0].p, &root, sizeof root, cudaMemcpyHostToDevice, s));
        TB_CUDA(cudaMemcpyAsync(f_off[0].p, off, sizeof off, cudaMemcpyHostToDevice, s));
        if (n_prims) iota_kernel<<<blocks(n_prims), kThreads, 0, s>>>(n_prims, refs[0].as<uint32_t>(), ref_f[0].as<uint32_t>());
        TB_CUDA(cudaStreamSynchronize(s)); // the two host arrays above die at scope exit
    }
    auto scan_u32 = [&](const uint32_t* in, uint32_t* out, size_t n) -> cudaError_t { // exclusive, n+1 outputs: out[n] = total
        size_t need = 0;
        cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, n + 1, s);
        if (e != cudaSuccess) return e;
        if ((e = temp.reserve(need, 0, s)) != cudaSuccess) return e;
        return cub::DeviceScan::ExclusiveSum(temp.p, need, in, out, n + 1, s);
    };
    const bool debug = std::getenv("G19_DEBUG_TREE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    while (a.F > 0) {
        const int nxt = cur ^ 1;
        const size_t F = a.F, R = a.R;
        TB_CUDA(mask.reserve(R + 1, 0, s));
        TB_CUDA(pos.reserve(sizeof(Lane8) * (R + 1), 0, s));
        TB_CUDA(child_cnt.reserve(sizeof(uint32_t) * 8 * F, 0, s));
        TB_CUDA(leaf_len.reserve(sizeof(uint32_t) * (F + 1), 0, s));
        TB_CUDA(interior.reserve(sizeof(uint32_t) * (F + 1), 0, s));
        TB_CUDA(leaf_off.reserve(sizeof(uint32_t) * (F + 1), 0, s));
        TB_CUDA(int_rank.reserve(sizeof(uint32_t) * (F + 1), 0, s));
        TB_CUDA(cudaMemsetAsync(child_cnt.p, 0, sizeof(uint32_t) * 8 * F, s));
        TB_CUDA(cudaMemsetAsync(leaf_len.p, 0, sizeof(uint32_t) * (F + 1), s));   // the extra element feeds the
        TB_CUDA(cudaMemsetAsync(interior.p, 0, sizeof(uint32_t) * (F + 1), s));   // "total" slot of the scans
        if (R) classify_kernel<<<blocks(R), kThreads, 0, s>>>(a, fr[cur].as<Frontier>(), f_off[cur].as<uint32_t>(), refs[cur].as<uint32_t>(),
                                                               ref_f[cur].as<uint32_t>(), mask.as<uint8_t>(), child_cnt.as<uint32_t>());
        decide_kernel<<<blocks(F), kThreads, 0, s>>>(a, f_off[cur].as<uint32_t>(), child_cnt.as<uint32_t>(), leaf_len.as<uint32_t>(),
                                                      interior.as<uint32_t>());
        TB_CUDA(scan_u32(leaf_len.as<uint32_t>(), leaf_off.as<uint32_t>(), F));
        TB_CUDA(scan_u32(interior.as<uint32_t>(), int_rank.as<uint32_t>(), F));
        uint32_t tot[2] = {0, 0}; // leaf references, interior nodes of this level
        TB_CUDA(cudaMemcpyAsync(&tot[0], leaf_off.as<uint32_t>() + F, 4, cudaMemcpyDeviceToHost, s));
        TB_CUDA(cudaMemcpyAsync(&tot[1], int_rank.as<uint32_t>() + F, 4, cudaMemcpyDeviceToHost, s));
        TB_CUDA(cudaStreamSynchronize(s));
        const size_t L = tot[0], I = tot[1], Fn = 8 * I;
        if (size_t(nodes_size) + Fn >= 0x7fffffffull || size_t(index_size) + L >= 0x7fffffffull) {
            err = "tree build: more than 2^31 nodes or leaf references";
            for (Buf* b : all) b->release();
            return G19_ERR_LIMIT;
        }
        TB_CUDA(nodes.reserve(sizeof(PathNodeD) * (size_t(nodes_size) + Fn), sizeof(PathNodeD) * size_t(nodes_size), s));
        TB_CUDA(index.reserve(sizeof(uint32_t) * (size_t(index_size) + L + 4), sizeof(uint32_t) * size_t(index_size), s));
        TB_CUDA(fr[nxt].reserve(sizeof(Frontier) * (Fn + 1), 0, s));
        TB_CUDA(next_cnt.reserve(sizeof(uint32_t) * (Fn + 1), 0, s));
        TB_CUDA(f_off[nxt].reserve(sizeof(uint32_t) * (Fn + 2), 0, s));
        TB_CUDA(cudaMemsetAsync(next_cnt.p, 0, sizeof(uint32_t) * (Fn + 1), s));
        emit_kernel<<<blocks(F), kThreads, 0, s>>>(a, fr[cur].as<Frontier>(), f_off[cur].as<uint32_t>(), interior.as<uint32_t>(),
                                                    int_rank.as<uint32_t>(), leaf_off.as<uint32_t>(), child_cnt.as<uint32_t>(),
                                                    nodes_size, index_size, nodes.as<PathNodeD>(), fr[nxt].as<Frontier>(),
                                                    next_cnt.as<uint32_t>());
        TB_CUDA(scan_u32(next_cnt.as<uint32_t>(), f_off[nxt].as<uint32_t>(), Fn));
        uint32_t Rn = 0;
        TB_CUDA(cudaMemcpyAsync(&Rn, f_off[nxt].as<uint32_t>() + Fn, 4, cudaMemcpyDeviceToHost, s));
        if (R) { // rank of every ref among its node's refs, per child: one exclusive scan of 8-lane counters
            auto lanes = thrust::make_transform_iterator(static_cast<const uint8_t*>(mask.as<uint8_t>()), MaskToLanes());
            size_t need = 0;
            Lane8 zero = {};
            TB_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, need, lanes, pos.as<Lane8>(), Lane8Add(), zero, R, s));
            TB_CUDA(temp.reserve(need, 0, s));
            TB_CUDA(cub::DeviceScan::ExclusiveScan(temp.p, need, lanes, pos.as<Lane8>(), Lane8Add(), zero, R, s));
        }
        TB_CUDA(cudaStreamSynchronize(s));
        TB_CUDA(refs[nxt].reserve(sizeof(uint32_t) * (size_t(Rn) + 1), 0, s));
        TB_CUDA(ref_f[nxt].reserve(sizeof(uint32_t) * (size_t(Rn) + 1), 0, s));
        if (R) scatter_kernel<<<blocks(R), kThreads, 0, s>>>(a, f_off[cur].as<uint32_t>(), refs[cur].as<uint32_t>(), ref_f[cur].as<uint32_t>(),
                                                              mask.as<uint8_t>(), pos.as<Lane8>(), interior.as<uint32_t>(),
                                                              int_rank.as<uint32_t>(), leaf_off.as<uint32_t>(), f_off[nxt].as<uint32_t>(),
                                                              index_size, index.as<uint32_t>(), refs[nxt].as<uint32_t>(),
                                                              ref_f[nxt].as<uint32_t>());
        TB_CUDA(cudaGetLastError());
        if (debug) {
            cudaStreamSynchronize(s);
            const auto t_now = std::chrono::steady_clock::now();
            std::fprintf(stderr, "[g19] tree level %2d: %9zu nodes, %9zu refs -> %zu leaf refs, %zu interior, %u refs below  (%.2f ms)\n",
                         a.level, F, R, L, I, Rn, std::chrono::duration<double, std::milli>(t_now - t_prev).count());
            t_prev = t_now;
        }
        if (I) depth = a.level + 1;
        nodes_size += uint32_t(Fn);
        index_size += uint32_t(L);
        a.F = uint32_t(Fn);
        a.R = Rn;
        a.level += 1;
        cur = nxt;
    }
    // the caller owns plain allocations of the exact size (the growable arrays go back to the pool)
    void *out_nodes = nullptr, *out_index = nullptr;
    TB_CUDA(cudaMalloc(&out_nodes, sizeof(PathNodeD) * size_t(nodes_size) + 256));
    cudaError_t oe = cudaMalloc(&out_index, sizeof(uint32_t) * size_t(index_size) + 256);
    if (oe == cudaSuccess) oe = cudaMemcpyAsync(out_nodes, nodes.p, sizeof(PathNodeD) * size_t(nodes_size), cudaMemcpyDeviceToDevice, s);
    if (oe == cudaSuccess && index_size) oe = cudaMemcpyAsync(out_index, index.p, sizeof(uint32_t) * size_t(index_size), cudaMemcpyDeviceToDevice, s);
    if (oe == cudaSuccess) oe = cudaStreamSynchronize(s);
    if (oe != cudaSuccess) {
        cudaFree(out_nodes);
        cudaFree(out_index);
        TB_CUDA(oe);
    }
    *d_nodes = static_cast<PathNodeD*>(out_nodes);
    *d_index = static_cast<uint32_t*>(out_index);
    *n_nodes = nodes_size;
    *n_index = index_size;
    *tree_depth = depth;
    for (Buf* b : all) b->release();
    return G19_OK;
#undef TB_CUDA
}


// ---- top table: direct index over the first levels of the octree -----------------------------------------------
// The walk's descents are a chain of DEPENDENT node loads, one per level (ncu, profiles/r02b: 23 % of trace_kernel's
// stall samples sit on the instruction behind that load, 7 of 32 lanes active). For every cell of level `top_level`
// this table holds the record of the deepest node on the way to it -- the cell's own node, or the leaf / empty node
// of a coarser level that covers it, with that level in bits 24-27 of the count word -- so a descent starts with
// ONE load instead of top_level of them.
__global__ void top_table_kernel(const PathNodeD* __restrict__ nodes, int top_level, uint2* __restrict__ table) {
    const uint32_t n = 1u << top_level;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n * n) return;
    const uint32_t x = i & (n - 1u), y = (i >> top_level) & (n - 1u), z = i >> (2 * top_level);
    uint2 rec = make_uint2(nodes[0].first, nodes[0].count);
    int level = 0;
    while (level < top_level && !(rec.y & kLeafBit)) {
        const int sh = top_level - 1 - level;
        const uint32_t c = ((x >> sh) & 1u) | (((y >> sh) & 1u) << 1) | (((z >> sh) & 1u) << 2);
        const PathNodeD nd = nodes[rec.x + c];
        rec = make_uint2(nd.first, nd.count);
        ++level;
    }
    if (rec.y & kLeafBit) rec.y |= uint32_t(level) << 24; // leaf counts stay far below 2^24
    table[i] = rec;
}

int path_build_top_table(const PathNodeD* d_nodes, int top_level, uint2* d_table, cudaStream_t s) {
    const uint32_t cells = 1u << (3 * top_level);
    top_table_kernel<<<(cells + 255) / 256, 256, 0, s>>>(d_nodes, top_level, d_table);
    return cudaGetLastError() == cudaSuccess ? G19_OK : G19_ERR_CUDA;
}

} // namespace g19
