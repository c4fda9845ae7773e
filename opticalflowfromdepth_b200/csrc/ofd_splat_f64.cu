// ofd_splat_f64.cu — the float64 instantiation of fw_cuda.forward_warping (AT_DISPATCH_FLOATING_TYPES also generates
// forward_warping_cuda_kernel<double>, alt_cuda/fw_cuda_kernel.cu:70; the reference's fw.py never uses it).
// A 64-bit ordered depth does not leave room for the source id in one 64-bit key, so the z-buffer is two planes
// (16 B/px of workspace): plane A takes atomicMin of the ordered depth, plane B atomicMin of the raster id among the
// depth-minimal sources; the gather reads both and re-arms both.  Same winner rule as the float path.
#include "ofd_common.cuh"

namespace ofd {

constexpr u64 HI64_NOWIN = 0xFFFFFFFFFFFFFFFEull;

__device__ __forceinline__ u64 depth_hi64(double d) {
    if (!(d < 1000.0)) return HI64_NOWIN;
    u64 b = (u64)__double_as_longlong(d);
    if (b == 0x8000000000000000ull) b = 0ull;
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ uint32_t target_f64(double x, double y, int H, int W) {
    if (!(x > -1.0 && x < (double)W && y > -1.0 && y < (double)H)) return T_DROPPED;
    return (uint32_t)((int)y * W + (int)x);
}

__global__ void __launch_bounds__(256) ztest64_depth_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                           const double* __restrict__ depth, u64* __restrict__ ka, size_t hw,
                                                           int H, int W, uint64_t* __restrict__ counters) {
    const int b = blockIdx.y;
    unsigned dropped = 0;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        const size_t q = (size_t)b * hw + p;
        const uint32_t t = target_f64(sx[q], sy[q], H, W);
        if (t == T_DROPPED)
            dropped++;
        else
            atomicMin(ka + (size_t)b * hw + t, depth_hi64(depth[q]));
    }
    if (counters) {
        // the grid-stride loop can leave some lanes without work: make the warp converge before the reduction
        __syncwarp();
        warp_count(counters, OFD_CNT_DROPPED, dropped);
    }
}

__global__ void __launch_bounds__(256) ztest64_index_kernel(const double* __restrict__ sx, const double* __restrict__ sy,
                                                           const double* __restrict__ depth, const u64* __restrict__ ka,
                                                           u64* __restrict__ kb, size_t hw, int H, int W) {
    const int b = blockIdx.y;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        const size_t q = (size_t)b * hw + p;
        const uint32_t t = target_f64(sx[q], sy[q], H, W);
        if (t != T_DROPPED && ka[(size_t)b * hw + t] == depth_hi64(depth[q])) atomicMin(kb + (size_t)b * hw + t, (u64)p);
    }
}

__global__ void __launch_bounds__(256) gather64_kernel(const double* __restrict__ obj, u64* __restrict__ ka, u64* __restrict__ kb,
                                                      size_t hw, int C, double* __restrict__ out, double* __restrict__ valid,
                                                      double* __restrict__ collision, int32_t* __restrict__ winner,
                                                      uint64_t* __restrict__ counters) {
    const int b = blockIdx.y;
    unsigned n_hit = 0, n_col = 0, n_px = 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < hw; t += (size_t)gridDim.x * blockDim.x) {
        const size_t q = (size_t)b * hw + t;
        const u64 a = ka[q], src = kb[q];
        const bool hit = a != KEY_UNTOUCHED;
        const bool win = a < HI64_NOWIN;
        for (int c = 0; c < C; ++c)
            out[((size_t)b * C + c) * hw + t] = win ? obj[((size_t)b * C + c) * hw + src] : 0.0;
        valid[q] = hit ? 1.0 : 0.0;
        collision[q] = (hit && !win) ? 1.0 : 0.0;
        if (winner) winner[q] = win ? (int32_t)src : (hit ? -2 : -1);
        ka[q] = KEY_UNTOUCHED;
        kb[q] = KEY_UNTOUCHED;
        n_px++, n_hit += hit, n_col += (hit && !win);
    }
    if (counters) {
        __syncwarp();
        warp_count(counters, OFD_CNT_HIT, n_hit);
        warp_count(counters, OFD_CNT_HOLE, n_px - n_hit);
        warp_count(counters, OFD_CNT_COLLISION, n_col);
    }
}

int splat_targets_f64(const char* fn, const double* obj, const double* sy, const double* sx, const double* depth, int B, int C,
                      int H, int W, double* out, double* valid, double* collision, int32_t* winner, uint64_t* counters,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t hw = (size_t)H * W;
    if (ws_bytes < 2 * (size_t)B * hw * sizeof(u64))
        return fail(OFD_E_WORKSPACE, "%s: float64 needs two key planes = %zu bytes of workspace (got %zu)", fn,
                    2 * (size_t)B * hw * sizeof(u64), ws_bytes);
    if (B > 65535) return fail(OFD_E_SHAPE, "%s: float64 path supports B <= 65535", fn);
    u64* ka = (u64*)ws;
    u64* kb = ka + (size_t)B * hw;
    size_t gx = (hw + 255) / 256;
    if (gx > 148 * 16) gx = 148 * 16;
    dim3 grid((unsigned)gx, B);
    ztest64_depth_kernel<<<grid, 256, 0, st>>>(sx, sy, depth, ka, hw, H, W, counters);
    ztest64_index_kernel<<<grid, 256, 0, st>>>(sx, sy, depth, ka, kb, hw, H, W);
    gather64_kernel<<<grid, 256, 0, st>>>(obj, ka, kb, hw, C, out, valid, collision, winner, counters);
    return check_launch(fn);
}

}  // namespace ofd
