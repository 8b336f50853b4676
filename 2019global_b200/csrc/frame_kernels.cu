// frame_kernels.cu -- the framebuffer "gather" of SURVEY.md 8(e) fused into the resolve step.
//
// Multi-GPU rendering shards image tiles over ranks (one process per GPU); the only exchange is
// bringing every rank's pixels to rank 0. Instead of staging a compact tile array, sending it
// with NCCL and scattering it on rank 0, the kernel that turns accumulated radiance into pixels
// stores them STRAIGHT INTO RANK 0's FRAME through a peer mapping (CUDA IPC over NVLink 5 /
// NVSwitch): resolve + untile + gather are one kernel, no collective call, no staging copy.
//
//   resolve_to_frame   accum / spp -> float radiance + truncated RGB888 (Image::setPixel, reference
//                      include/image.h:14-16) at the pixel's place in the (possibly remote) frame;
//                      4 pixels per thread: 3 x 16 B radiance stores + 3 x 4 B RGB stores
//   frame_signal       one thread: system-scope fence, then arrived += 1 in the frame owner's memory
//   frame_wait         one thread on the owner: spin (acquire, system scope) until arrived reaches
//                      epoch * world -- everything enqueued behind it sees the whole frame
//   frame_release      owner: consumed = epoch; peers' frame_acquire spins on it before they overwrite
//
// The reference has no counterpart (single process, single thread: raytracer.h:23-87).
#include "kernels.h"

namespace g19 {
namespace {

constexpr int kThreads = 256;
constexpr unsigned long long kSpinLimitNs = 5000000000ull; // a lost peer must not hang the GPU

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t quantise(float v) { // truncation, as Image::setPixel
    float cl = fminf(fmaxf(v, 0.0f), 1.0f);
    return (uint32_t)(int)(255.0f * cl);
}

__global__ void __launch_bounds__(kThreads) resolve_to_frame_kernel(TileMap map, const float* __restrict__ accum, int spp,
                                                                    uint8_t* rgb_f, float* rad_f) {
    const uint32_t npix = (uint32_t)map.n_local_pix;
    const uint32_t lp = 4u * (blockIdx.x * kThreads + threadIdx.x); // 4 consecutive pixels of one tile row
    if (lp >= npix) return;
    const uint32_t lt = lp >> 10, in = lp & 1023u;
    const uint32_t t = lt * (uint32_t)map.world + (uint32_t)map.rank;
    const uint32_t ty = t / (uint32_t)map.tiles_x, tx = t - ty * (uint32_t)map.tiles_x;
    const int x = int(tx * kTile + (in & 31u)), y = int(ty * kTile + (in >> 5));
    if (y >= map.h || x >= map.w) return;
    const float4 a0 = *reinterpret_cast<const float4*>(accum + lp);
    const float4 a1 = *reinterpret_cast<const float4*>(accum + (size_t)npix + lp);
    const float4 a2 = *reinterpret_cast<const float4*>(accum + 2 * (size_t)npix + lp);
    const float s = float(spp);
    float r[4] = {__fdiv_rn(a0.x, s), __fdiv_rn(a0.y, s), __fdiv_rn(a0.z, s), __fdiv_rn(a0.w, s)};
    float g[4] = {__fdiv_rn(a1.x, s), __fdiv_rn(a1.y, s), __fdiv_rn(a1.z, s), __fdiv_rn(a1.w, s)};
    float b[4] = {__fdiv_rn(a2.x, s), __fdiv_rn(a2.y, s), __fdiv_rn(a2.z, s), __fdiv_rn(a2.w, s)};
    const size_t gpx = (size_t)y * map.w + x;
    if (x + 3 < map.w && (map.w & 3) == 0) { // aligned: (y*w + x) is a multiple of 4
        if (rad_f) {
            float4* o = reinterpret_cast<float4*>(rad_f + 3 * gpx);
            o[0] = make_float4(r[0], g[0], b[0], r[1]);
            o[1] = make_float4(g[1], b[1], r[2], g[2]);
            o[2] = make_float4(b[2], r[3], g[3], b[3]);
        }
        if (rgb_f) {
            uint32_t q[12];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                q[3 * k] = quantise(r[k]);
                q[3 * k + 1] = quantise(g[k]);
                q[3 * k + 2] = quantise(b[k]);
            }
            uint32_t* o = reinterpret_cast<uint32_t*>(rgb_f + 3 * gpx);
            o[0] = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
            o[1] = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
            o[2] = q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24);
        }
    } else { // ragged right edge or unaligned width
        for (int k = 0; k < 4 && x + k < map.w; ++k) {
            if (rad_f) {
                rad_f[3 * (gpx + k)] = r[k];
                rad_f[3 * (gpx + k) + 1] = g[k];
                rad_f[3 * (gpx + k) + 2] = b[k];
            }
            if (rgb_f) {
                rgb_f[3 * (gpx + k)] = (uint8_t)quantise(r[k]);
                rgb_f[3 * (gpx + k) + 1] = (uint8_t)quantise(g[k]);
                rgb_f[3 * (gpx + k) + 2] = (uint8_t)quantise(b[k]);
            }
        }
    }
}

// flags[0] = arrived (ranks that finished writing, summed over all frames so far)
// flags[32] = consumed (last epoch the owner is done reading), flags[64] = spin timeouts
__global__ void frame_signal_kernel(unsigned* flags) {
    __threadfence_system(); // the stores of every earlier kernel in this stream are performed; order them before the signal
    atomicAdd_system(flags, 1u);
}

// A spin that gives up poisons the frame: it counts the timeout in the owner's flags AND raises the sticky
// status word in this process's mapped host memory, which every later g19_frame_* call refuses on
// (G19_ERR_TIMEOUT) -- an incomplete frame is never handed out as G19_OK.
__device__ __forceinline__ void frame_give_up(unsigned* flags, unsigned* status) {
    atomicAdd_system(flags + 64, 1u);
    if (status) {
        *reinterpret_cast<volatile unsigned*>(status) = 1u;
        __threadfence_system();
    }
}

__global__ void frame_wait_kernel(unsigned* flags, unsigned target, unsigned* status) {
    const unsigned long long t0 = now_ns();
    while (int(ld_acquire_sys(flags) - target) < 0) {
        __nanosleep(200);
        if (now_ns() - t0 > kSpinLimitNs) {
            frame_give_up(flags, status);
            break;
        }
    }
}

__global__ void frame_release_kernel(unsigned* flags, unsigned epoch) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags + 32), "r"(epoch) : "memory");
}

__global__ void frame_acquire_kernel(unsigned* flags, unsigned need_consumed, unsigned* status) {
    const unsigned long long t0 = now_ns();
    while (int(ld_acquire_sys(flags + 32) - need_consumed) < 0) {
        __nanosleep(500);
        if (now_ns() - t0 > kSpinLimitNs) {
            frame_give_up(flags, status);
            break;
        }
    }
}

} // namespace

void launch_resolve_to_frame(const TileMap& map, const float* accum, int spp, uint8_t* rgb_frame, float* rad_frame,
                             cudaStream_t s) {
    if (map.n_local_pix == 0) return;
    const int quads = map.n_local_pix / 4;
    resolve_to_frame_kernel<<<(quads + kThreads - 1) / kThreads, kThreads, 0, s>>>(map, accum, spp, rgb_frame, rad_frame);
}
void launch_frame_signal(unsigned* flags, cudaStream_t s) { frame_signal_kernel<<<1, 1, 0, s>>>(flags); }
void launch_frame_wait(unsigned* flags, unsigned target, unsigned* status, cudaStream_t s) { frame_wait_kernel<<<1, 1, 0, s>>>(flags, target, status); }
void launch_frame_release(unsigned* flags, unsigned epoch, cudaStream_t s) { frame_release_kernel<<<1, 1, 0, s>>>(flags, epoch); }
void launch_frame_acquire(unsigned* flags, unsigned need, unsigned* status, cudaStream_t s) { frame_acquire_kernel<<<1, 1, 0, s>>>(flags, need, status); }

} // namespace g19
