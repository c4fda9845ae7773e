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
#ifndef OFD_RUNMIN_MAX_HEADS
#define OFD_RUNMIN_MAX_HEADS 32
#endif
__device__ __forceinline__ bool warp_run_min(uint32_t t, u64& key, int lane) {
    const unsigned full = 0xFFFFFFFFu;
    uint32_t t_prev = __shfl_up_sync(full, t, 1);
    bool head = (lane == 0) || (t != t_prev);
    unsigned heads = __ballot_sync(full, head);
    // Experiment knob (default off): skip the pre-reduction when the warp has more than OFD_RUNMIN_MAX_HEADS distinct runs
    // and let every lane issue its own atomic.  MEASURED slower on B200 (profiles/r1/tune_runmin.txt: FW C=6 600 -> 626 us
    // at 16, 670 us at 8): duplicate same-address atomics cost more in L2 than the 42 shuffle/select instructions here.
    if (OFD_RUNMIN_MAX_HEADS < 32 && __popc(heads) > OFD_RUNMIN_MAX_HEADS) return t != T_DROPPED;
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

// ---- the z-buffer plane as stored in the workspace -------------------------------------------------------------------
// OFD_KEY32 = 0: one packed 64-bit key per target, one fire-and-forget REDG.MIN.64 per source.
// OFD_KEY32 = 1: only the 32-bit source id is stored (4 B per target: half the key traffic of z-test RMW, gather read and
//                re-arm); the ordering key (ordered depth, id) of the current holder is re-derived from the read-only
//                depth plane, and the minimum is taken by a compare-and-swap loop.  Same winner: the loop only ever
//                replaces a holder by a source with a strictly smaller (depth, id) key, so the final holder is the
//                minimum, whatever the order the SMs run in.
// MEASURED (B200, profiles/r1/tune_key32.txt; 128 x 480x640): bit-identical (all GPU parity tests pass with either), but
// the returning CAS costs more than the 16 B/px it saves: FW C=6 752 us vs 600 us, reproject pair 1071 vs 732 us.  The
// packed 64-bit key with a fire-and-forget RED stays the default; -DOFD_KEY32=1 rebuilds the experiment.
#ifndef OFD_KEY32
#define OFD_KEY32 0
#endif
#if OFD_KEY32
typedef uint32_t zkey_t;
constexpr zkey_t ZKEY_EMPTY = 0xFFFFFFFFu;  // no source reached this target (the 0xFF workspace pattern)
constexpr zkey_t ZKEY_NOWIN = 0xFFFFFFFEu;  // reached only by sources with !(depth < 1000)

__device__ __forceinline__ void zkey_min(zkey_t* __restrict__ kp, const float* __restrict__ dp, uint32_t t, u64 key) {
    const uint32_t src = (uint32_t)key;
    if ((uint32_t)(key >> 32) >= HI_NOWIN) {
        atomicCAS(kp + t, ZKEY_EMPTY, ZKEY_NOWIN);  // marks "hit"; any real source overrides it
        return;
    }
    zkey_t old = atomicCAS(kp + t, ZKEY_EMPTY, src);
    while (old != ZKEY_EMPTY) {
        if (old != ZKEY_NOWIN && make_key(depth_hi(__ldg(dp + old)), old) < key) break;  // the holder wins
        const zkey_t seen = atomicCAS(kp + t, old, src);
        if (seen == old) break;
        old = seen;
    }
}
__device__ __forceinline__ bool zkey_hit(zkey_t k) { return k != ZKEY_EMPTY; }
__device__ __forceinline__ bool zkey_win(zkey_t k) { return k < ZKEY_NOWIN; }
__device__ __forceinline__ uint32_t zkey_src(zkey_t k) { return k; }
// ordered depth of the winner (tie census)
__device__ __forceinline__ uint32_t zkey_winner_hi(zkey_t k, const float* __restrict__ dp) {
    return zkey_win(k) ? depth_hi(__ldg(dp + k)) : HI_NOWIN;
}
#else
typedef u64 zkey_t;
constexpr zkey_t ZKEY_EMPTY = KEY_UNTOUCHED;
__device__ __forceinline__ void zkey_min(zkey_t* __restrict__ kp, const float* __restrict__, uint32_t t, u64 key) {
    atomicMin(kp + t, key);
}
__device__ __forceinline__ bool zkey_hit(zkey_t k) { return k != KEY_UNTOUCHED; }
__device__ __forceinline__ bool zkey_win(zkey_t k) { return (uint32_t)(k >> 32) < HI_NOWIN; }
__device__ __forceinline__ uint32_t zkey_src(zkey_t k) { return (uint32_t)k; }
__device__ __forceinline__ uint32_t zkey_winner_hi(zkey_t k, const float* __restrict__) { return (uint32_t)(k >> 32); }
#endif

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

// ---- 6-DoF reprojection flow -----------------------------------------------------------------------------
struct Cam {  // 21 floats per frame: invK3 row-major (9), P = (K T)[:3,:] row-major (12)
    float k[9];
    float p[12];
};

template <typename DT>
__device__ __forceinline__ void reproject_px(const Cam& cam, DT depth, int i, int j, int H, int W, float eps,
                                             float& fx, float& fy) {
    const float x = (float)i, y = (float)j;
    // geometry.py:38  cam_points = inv_K[:3,:3] @ (x, y, 1)
    float ray[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(cam.k[3 * r + 0], x);
        acc = __fmaf_rn(cam.k[3 * r + 1], y, acc);
        acc = __fmaf_rn(cam.k[3 * r + 2], 1.0f, acc);
        ray[r] = acc;
    }
    // geometry.py:39-40  depth * cam_points in the depth dtype, then .type(float32)
    float X[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) X[r] = (float)(depth * (DT)ray[r]);
    // geometry.py:59  P @ (X, 1)
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(cam.p[4 * r + 0], X[0]);
        acc = __fmaf_rn(cam.p[4 * r + 1], X[1], acc);
        acc = __fmaf_rn(cam.p[4 * r + 2], X[2], acc);
        acc = __fmaf_rn(cam.p[4 * r + 3], 1.0f, acc);
        c[r] = acc;
    }
    // geometry.py:61  / (z + eps)
    const float den = __fadd_rn(c[2], eps);
    float u = __fdiv_rn(c[0], den);
    float v = __fdiv_rn(c[1], den);
    // geometry.py:64-66  /= (w-1), /= (h-1), (p - 0.5) * 2
    u = __fmul_rn(__fsub_rn(__fdiv_rn(u, (float)(W - 1)), 0.5f), 2.0f);
    v = __fmul_rn(__fsub_rn(__fdiv_rn(v, (float)(H - 1)), 0.5f), 2.0f);
    // preprocess.py:284-286  (p + 1) / 2, *= (w-1), *= (h-1)
    // (x / 2 is evaluated as x * 0.5f: a power-of-two scaling, bit-identical to the IEEE division)
    u = __fmul_rn(__fmul_rn(__fadd_rn(u, 1.0f), 0.5f), (float)(W - 1));
    v = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), (float)(H - 1));
    // preprocess.py:288-291  flow = p1 - p0
    fx = __fsub_rn(u, x);
    fy = __fsub_rn(v, y);
}

}  // namespace ofd

// ---- host side error plumbing (ofd_abi.cu owns the storage) ----------------------------------------------
namespace ofd {
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
// memoised per (device, kernel[, threads, smem]): dynamic shared-memory opt-in; SM count and resident CTAs per SM (ofd_abi.cu)
int ensure_dynamic_smem(const char* fn, const void* kern, size_t smem);
int launch_plan(const char* fn, const void* kern, int threads, size_t smem, int* sms, int* per_sm);
}  // namespace ofd
