// ofd_common.cuh — shared device helpers for libofd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ofd_b200.h"

namespace ofd {

typedef unsigned long long u64;

// ---- packed z-buffer key -------------------------------------------------------------------------------
// key = ordered(depth) << 32 | source raster id.   atomicMin over keys == (min depth, then min raster id),
// which is what the reference's serial raster loop with a strict '<' computes (fw_cuda_kernel.cu:28-36).
constexpr u64 KEY_UNTOUCHED = 0xFFFFFFFFFFFFFFFFull;  // no source reached this target -> valid = 0
constexpr uint32_t HI_NOWIN = 0xFFFFFFFEu;            // reached, but only by sources with !(depth < 1000)
constexpr uint32_t T_DROPPED = 0xFFFFFFFFu;           // source has no target (NaN flow / out of range)
constexpr float DLUT_INIT = 1000.0f;                  // fw_cuda_kernel.cu:58

// Order-preserving map float -> uint32 for depth < 1000 (never NaN here); -0.0 and +0.0 tie, as '<' says.
__device__ __forceinline__ uint32_t depth_hi(float d) {
    if (!(d < DLUT_INIT)) return HI_NOWIN;  // also catches NaN
    uint32_t b = __float_as_uint(d);
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ u64 make_key(uint32_t hi, uint32_t src) { return ((u64)hi << 32) | (u64)src; }

// Target of a source under FW.forward (alt_cuda/fw.py:31,37-42), evaluated in T = promote(float32, flow dtype).
template <typename T>
__device__ __forceinline__ uint32_t fw_target(int i, int j, T fx, T fy, int H, int W) {
    T px = (T)(float)i + fx;
    T py = (T)(float)j + fy;
    if (px != px || py != py) return T_DROPPED;
    // torch.clamp(x, min, max) == min(max(x, min), max)
    px = px < (T)0 ? (T)0 : px;
    px = px > (T)(W - 1) ? (T)(W - 1) : px;
    py = py < (T)0 ? (T)0 : py;
    py = py > (T)(H - 1) ? (T)(H - 1) : py;
    return (uint32_t)((int)py * W + (int)px);  // .type(int64): truncation toward zero
}

// ---- warp run aggregation --------------------------------------------------------------------------------
// Lanes hold consecutive raster sources.  Consecutive lanes that hit the same target (clamped borders,
// compressed regions) are merged to one key before the L2 atomic: a segmented min over runs.
// Returns true when this lane must issue the atomic with `key`.
__device__ __forceinline__ bool warp_run_min(uint32_t t, u64& key, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    uint32_t t_prev = __shfl_up_sync(full, t, 1);
    bool head = (lane == 0) || (t != t_prev);
    unsigned heads = __ballot_sync(full, head);
    if (heads != full) {  // warp-uniform
        unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
        int run_end = above ? (__ffs(above) - 2) : 31;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 o = __shfl_down_sync(full, key, d);
            if (lane + d <= run_end && o < key) key = o;
        }
    }
    return head && (t != T_DROPPED);
}

__device__ __forceinline__ void key_min(u64* addr, u64 key) { atomicMin(addr, key); }

// utils.fix_warped_depth (utils.py:123-126)
__device__ __forceinline__ float fix_depth(float d) {
    if (d == 0.0f) d = 100.0f;
    if (d > 99.5f) d = 100.0f;
    return d;
}

// counters: warp-reduce then one atomic per warp
__device__ __forceinline__ void warp_count(uint64_t* counters, int slot, unsigned n_lane) {
    unsigned s = __reduce_add_sync(0xFFFFFFFFu, n_lane);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd((u64*)(counters + slot), (u64)s);
}

}  // namespace ofd

// ---- host side error plumbing (ofd_abi.cu owns the storage) ----------------------------------------------
namespace ofd {
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
}  // namespace ofd
