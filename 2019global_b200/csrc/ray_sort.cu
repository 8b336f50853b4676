// ray_sort.cu -- tree scenes: the ray queue of a bounce, reordered by origin cell before trace_kernel walks it.
//
// What the end-state capture of trace_kernel on the 1 M-triangle room showed (profiles/r02g_tree_kernels_room_ncu.json):
// 8.1 warps stalled on long-scoreboard per issue, L1 hit rate 44 %, 20 of 32 lanes active -- after the first diffuse
// bounce the 32 rays of a warp start in 32 different places of the scene, every one of them pulls its own nodes, leaf
// lists and primitive records through L1, and their walks have nothing in common. The bounce kernels therefore tag every
// ray they queue with a 16-bit key (shadow rays: 1 | 15-bit Morton code of the origin; continuation rays:
// 0 | 12-bit Morton code | direction octant), and one LSD radix sort (CUB one-sweep, two 8-bit passes, ~12 bytes of
// traffic per ray and pass) turns the keys into the order in which trace_kernel fetches the rays. The records themselves
// do not move: trace_kernel reads them through the permutation.
//
// The number of rays is known on the device only, so the sort always covers the queue's capacity: the keys are
// pre-set to 0xffff, the sort is stable, hence the first n entries of the permutation are exactly the n queued rays
// (valid ones by key, the producers' chunk padding last).
#include <cub/device/device_radix_sort.cuh>

#include "path.h"

namespace g19 {

namespace {
__global__ void iota_kernel(uint32_t* __restrict__ out, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}
} // namespace

size_t ray_sort_temp_bytes(size_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint16_t*>(nullptr), static_cast<uint16_t*>(nullptr),
                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), n, 0, 16);
    return bytes;
}

void launch_iota(uint32_t* out, size_t n, cudaStream_t s) {
    if (n) iota_kernel<<<1184, 256, 0, s>>>(out, uint32_t(n));
}

cudaError_t launch_ray_sort(void* temp, size_t temp_bytes, const uint16_t* keys_in, uint16_t* keys_out, const uint32_t* iota,
                            uint32_t* perm, size_t n, cudaStream_t s) {
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, iota, perm, n, 0, 16, s);
}

} // namespace g19
