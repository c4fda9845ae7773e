// ofd_flow.cu — per-pixel flow producers and depth helpers, fused versions of the torch op chains in
// preprocess.py:265-298 + geometry.py:37-67 (6-DoF reprojection flow), preprocess.py:24-105 (SpecialFlow),
// utils.py:102-116 (normalize_depth) and utils.py:123-126 (fix_warped_depth).
//
// Every arithmetic step is written with explicit round-to-nearest intrinsics in the reference's op order so that
// no FMA contraction changes a rounding; only the reference's small matmuls (K=3/4 dot products, whose
// accumulation order is a BLAS implementation detail) are evaluated as ascending-k FMA chains.
#include "ofd_common.cuh"

namespace ofd {

// ---- 6-DoF reprojection flow: Cam / reproject_px live in ofd_common.cuh (shared with the fused splat) ----
template <typename DT>
__global__ void __launch_bounds__(256) reproject_flow_kernel(const DT* __restrict__ depth, const Cam* __restrict__ cams,
                                                            float eps, int H, int W, float* __restrict__ flow) {
    const int b = blockIdx.z;
    const int j = blockIdx.y * 8 + threadIdx.y;
    const size_t hw = (size_t)H * W;
    __shared__ Cam cam;
    if (threadIdx.y == 0 && threadIdx.x < 21)
        reinterpret_cast<float*>(&cam)[threadIdx.x] = reinterpret_cast<const float*>(cams + b)[threadIdx.x];
    __syncthreads();
    if (j >= H) return;
    const DT* d = depth + (size_t)b * hw + (size_t)j * W;
    float* f = flow + (size_t)b * 2 * hw + (size_t)j * W;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = blockIdx.x * 128 + 32 * k + threadIdx.x;
        if (i < W) {
            float fx, fy;
            reproject_px<DT>(cam, d[i], i, j, H, W, eps, fx, fy);
            f[i] = fx;
            f[hw + i] = fy;
        }
    }
}


// ---- geometry.py class-level entry points (kept for drop-in use; the fused kernel above is the fast path) ---
// BackprojectDepth.forward (geometry.py:37-42): cam_points[b, 0:3, p] = float32(depth * (invK3 (x,y,1))), [b,3,p] = 1
template <typename DT>
__global__ void __launch_bounds__(256) backproject_kernel(const DT* __restrict__ depth, const float* __restrict__ invk,
                                                         int H, int W, float* __restrict__ pts) {
    const int b = blockIdx.z, j = blockIdx.y * 8 + threadIdx.y, i = blockIdx.x * 32 + threadIdx.x;
    if (j >= H || i >= W) return;
    const size_t hw = (size_t)H * W, p = (size_t)j * W + i;
    const float* k = invk + 9 * b;
    const float x = (float)i, y = (float)j;
    const DT d = depth[(size_t)b * hw + p];
    float* o = pts + (size_t)b * 4 * hw + p;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(__ldg(k + 3 * r), x);
        acc = __fmaf_rn(__ldg(k + 3 * r + 1), y, acc);
        acc = __fmaf_rn(__ldg(k + 3 * r + 2), 1.0f, acc);
        o[r * hw] = (float)(d * (DT)acc);
    }
    o[3 * hw] = 1.0f;
}

// Project3D.forward (geometry.py:56-67): pix[b,j,i,:] = ((P pts).xy / ((P pts).z + eps) / (w-1,h-1) - .5) * 2 ; z = (P pts).z
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ pts, const float* __restrict__ P, float eps,
                                                     int H, int W, float* __restrict__ pix, float* __restrict__ z) {
    const int b = blockIdx.z, j = blockIdx.y * 8 + threadIdx.y, i = blockIdx.x * 32 + threadIdx.x;
    if (j >= H || i >= W) return;
    const size_t hw = (size_t)H * W, p = (size_t)j * W + i;
    const float* q = pts + (size_t)b * 4 * hw + p;
    const float X0 = q[0], X1 = q[hw], X2 = q[2 * hw], X3 = q[3 * hw];
    const float* m = P + 12 * b;
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(__ldg(m + 4 * r), X0);
        acc = __fmaf_rn(__ldg(m + 4 * r + 1), X1, acc);
        acc = __fmaf_rn(__ldg(m + 4 * r + 2), X2, acc);
        acc = __fmaf_rn(__ldg(m + 4 * r + 3), X3, acc);
        c[r] = acc;
    }
    const float den = __fadd_rn(c[2], eps);
    float u = __fdiv_rn(__fdiv_rn(c[0], den), (float)(W - 1));
    float v = __fdiv_rn(__fdiv_rn(c[1], den), (float)(H - 1));
    float2 o;
    o.x = __fmul_rn(__fsub_rn(u, 0.5f), 2.0f);
    o.y = __fmul_rn(__fsub_rn(v, 0.5f), 2.0f);
    reinterpret_cast<float2*>(pix)[(size_t)b * hw + p] = o;
    z[(size_t)b * hw + p] = c[2];
}

// ---- SpecialFlow -----------------------------------------------------------------------------------------
// kind 5: vertical flip (preprocess.py:54, the branch the reference always takes);
// kind 6/7: p1 = (p0 - c) @ M + c and p_prev = (p0 - c) @ Mrev + c (rotate :74-76; shear :93-95 with c = 0).
struct SpecialParams {
    float cx, cy;
    float m[4];     // row-major 2x2, p @ M  ->  x' = x*m00 + y*m10, y' = x*m01 + y*m11
    float mrev[4];
    int use_center;
};

__device__ __forceinline__ void special_flow_px(int kind, const SpecialParams& sp, int H, int W, int i, int j,
                                                float* __restrict__ flow, float* __restrict__ back) {
    const size_t hw = (size_t)H * W, p = (size_t)j * W + i;
    const float x = (float)i, y = (float)j;
    if (kind == 5) {
        const float fy = __fsub_rn((float)(H - 1 - j), y);
        flow[p] = 0.0f;
        flow[hw + p] = fy;
        back[p] = 0.0f;
        back[hw + p] = fy;
        return;
    }
    float dx = x, dy = y;
    if (sp.use_center) {
        dx = __fsub_rn(x, sp.cx);
        dy = __fsub_rn(y, sp.cy);
    }
    float x1 = __fmaf_rn(dy, sp.m[2], __fmul_rn(dx, sp.m[0]));
    float y1 = __fmaf_rn(dy, sp.m[3], __fmul_rn(dx, sp.m[1]));
    float x0 = __fmaf_rn(dy, sp.mrev[2], __fmul_rn(dx, sp.mrev[0]));
    float y0 = __fmaf_rn(dy, sp.mrev[3], __fmul_rn(dx, sp.mrev[1]));
    if (sp.use_center) {
        x1 = __fadd_rn(x1, sp.cx), y1 = __fadd_rn(y1, sp.cy);
        x0 = __fadd_rn(x0, sp.cx), y0 = __fadd_rn(y0, sp.cy);
    }
    flow[p] = __fsub_rn(x1, x);
    flow[hw + p] = __fsub_rn(y1, y);
    back[p] = __fsub_rn(x0, x);
    back[hw + p] = __fsub_rn(y0, y);
}

__global__ void __launch_bounds__(256) special_flow_kernel(int kind, SpecialParams sp, int H, int W,
                                                          float* __restrict__ flow, float* __restrict__ back) {
    const int j = blockIdx.y * 8 + threadIdx.y;
    const int i = blockIdx.x * 32 + threadIdx.x;
    if (j >= H || i >= W) return;
    special_flow_px(kind, sp, H, W, i, j, flow, back);
}

// One special flow per sample of a batch (in-loop augmentation, BASELINE config 4): the per-sample kind and parameters
// travel in the kernel parameter block, SPECIAL_BATCH samples per launch.
constexpr int SPECIAL_BATCH = 48;
struct SpecialBatch {
    SpecialParams sp[SPECIAL_BATCH];
    int kind[SPECIAL_BATCH];
};

__global__ void __launch_bounds__(256) special_flow_batch_kernel(const __grid_constant__ SpecialBatch sb, int H, int W,
                                                                float* __restrict__ flow, float* __restrict__ back) {
    const int b = blockIdx.z;
    const int j = blockIdx.y * 8 + threadIdx.y;
    const int i = blockIdx.x * 32 + threadIdx.x;
    if (j >= H || i >= W) return;
    const size_t off = (size_t)b * 2 * H * W;
    special_flow_px(sb.kind[b], sb.sp[b], H, W, i, j, flow + off, back + off);
}

// ---- normalize_depth ---------------------------------------------------------------------------------------
template <typename DT>
struct Ord;
template <>
struct Ord<float> {
    typedef uint32_t U;
    static __device__ __forceinline__ U enc(float v) {
        uint32_t b = __float_as_uint(v);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    static __device__ __forceinline__ float dec(U u) {
        return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
    }
};
template <>
struct Ord<double> {
    typedef u64 U;
    static __device__ __forceinline__ U enc(double v) {
        u64 b = (u64)__double_as_longlong(v);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double dec(U u) {
        return __longlong_as_double((long long)((u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u));
    }
};

__global__ void minmax_init_kernel(u64* scratch, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        scratch[2 * b] = ~0ull;  // running min (ordered encoding)
        scratch[2 * b + 1] = 0ull;  // running max
    }
}

// utils.py:103-108: m1 = (d == 0 or d > 100) ? 100 : d ; min over m1 ; m2 = (m1 == 100) ? 0 : m1 ; max over m2
template <typename DT>
__global__ void __launch_bounds__(256) minmax_kernel(const DT* __restrict__ depth, size_t hw, u64* __restrict__ scratch) {
    typedef typename Ord<DT>::U U;
    const int b = blockIdx.y;
    const DT* d = depth + (size_t)b * hw;
    U mn = ~(U)0, mx = 0;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        DT v = d[p];
        if (v == (DT)0 || v > (DT)100) v = (DT)100;
        U e1 = Ord<DT>::enc(v);
        if (v == (DT)100) v = (DT)0;
        U e2 = Ord<DT>::enc(v);
        mn = e1 < mn ? e1 : mn;
        mx = e2 > mx ? e2 : mx;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        U a = __shfl_xor_sync(0xFFFFFFFFu, mn, o), c = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
        mn = a < mn ? a : mn;
        mx = c > mx ? c : mx;
    }
    __shared__ U smn[8], smx[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) smn[w] = mn, smx[w] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            mn = smn[k] < mn ? smn[k] : mn;
            mx = smx[k] > mx ? smx[k] : mx;
        }
        atomicMin(scratch + 2 * b, (u64)mn);
        atomicMax(scratch + 2 * b + 1, (u64)mx);
    }
}

// utils.py:109-110: out = (m2 - mn) * 98 / (mx - mn) + 1 ; out[out == image_of_zero] = 100
template <typename DT>
__global__ void __launch_bounds__(256) normalize_map_kernel(const DT* __restrict__ depth, size_t hw,
                                                           const u64* __restrict__ scratch, DT* __restrict__ out) {
    typedef typename Ord<DT>::U U;
    const int b = blockIdx.y;
    const DT mn = Ord<DT>::dec((U)scratch[2 * b]);
    const DT mx = Ord<DT>::dec((U)scratch[2 * b + 1]);
    const DT range = mx - mn;
    const DT zero_img = (((DT)0 - mn) * (DT)98) / range + (DT)1;
    const DT* d = depth + (size_t)b * hw;
    DT* o = out + (size_t)b * hw;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        DT v = d[p];
        if (v == (DT)0 || v > (DT)100) v = (DT)100;
        if (v == (DT)100) v = (DT)0;
        DT r = ((v - mn) * (DT)98) / range + (DT)1;
        if (r == zero_img) r = (DT)100;
        o[p] = r;
    }
}

// Ragged variants (BASELINE config 2: mixed-resolution frames packed back to back): image blockIdx.y is `count[y]` elements
// at element offset `offset[y]`; the arithmetic is the per-frame kernels' above.
constexpr int RAGGED_MAX = 128;
struct RaggedTable {
    unsigned long long offset[RAGGED_MAX];
    unsigned long long count[RAGGED_MAX];
};

template <typename DT>
__global__ void __launch_bounds__(256) minmax_ragged_kernel(const DT* __restrict__ depth, const __grid_constant__ RaggedTable tab,
                                                           u64* __restrict__ scratch) {
    typedef typename Ord<DT>::U U;
    const int b = blockIdx.y;
    const DT* d = depth + tab.offset[b];
    const size_t hw = tab.count[b];
    U mn = ~(U)0, mx = 0;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        DT v = d[p];
        if (v == (DT)0 || v > (DT)100) v = (DT)100;
        U e1 = Ord<DT>::enc(v);
        if (v == (DT)100) v = (DT)0;
        U e2 = Ord<DT>::enc(v);
        mn = e1 < mn ? e1 : mn;
        mx = e2 > mx ? e2 : mx;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        U a = __shfl_xor_sync(0xFFFFFFFFu, mn, o), c = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
        mn = a < mn ? a : mn;
        mx = c > mx ? c : mx;
    }
    __shared__ U smn[8], smx[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) smn[w] = mn, smx[w] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            mn = smn[k] < mn ? smn[k] : mn;
            mx = smx[k] > mx ? smx[k] : mx;
        }
        atomicMin(scratch + 2 * b, (u64)mn);
        atomicMax(scratch + 2 * b + 1, (u64)mx);
    }
}

template <typename DT>
__global__ void __launch_bounds__(256) normalize_map_ragged_kernel(const DT* __restrict__ depth, const __grid_constant__ RaggedTable tab,
                                                                  const u64* __restrict__ scratch, DT* __restrict__ out) {
    typedef typename Ord<DT>::U U;
    const int b = blockIdx.y;
    const DT mn = Ord<DT>::dec((U)scratch[2 * b]);
    const DT mx = Ord<DT>::dec((U)scratch[2 * b + 1]);
    const DT range = mx - mn;
    const DT zero_img = (((DT)0 - mn) * (DT)98) / range + (DT)1;
    const DT* d = depth + tab.offset[b];
    DT* o = out + tab.offset[b];
    const size_t hw = tab.count[b];
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        DT v = d[p];
        if (v == (DT)0 || v > (DT)100) v = (DT)100;
        if (v == (DT)100) v = (DT)0;
        DT r = ((v - mn) * (DT)98) / range + (DT)1;
        if (r == zero_img) r = (DT)100;
        o[p] = r;
    }
}

__global__ void __launch_bounds__(256) fix_depth_kernel(float* __restrict__ d, size_t n) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        d[p] = fix_depth(d[p]);
}

// ---- utils.inpaint hole-mask logic (utils.py:137-149) ----------------------------------------------
// M = (valid != collision); M' = 3x3 dilation of M (cv2.dilate, out-of-image taps ignored); P = (M' == M);
// H' = valid * P; mask = 1 - H'  (uint8: the mask handed to cv2.inpaint)
__global__ void __launch_bounds__(256) inpaint_mask_kernel(const float* __restrict__ valid, const float* __restrict__ collision,
                                                          int H, int W, unsigned char* __restrict__ mask) {
    const int b = blockIdx.z, j = blockIdx.y * 8 + threadIdx.y, i = blockIdx.x * 32 + threadIdx.x;
    if (j >= H || i >= W) return;
    const size_t base = (size_t)b * H * W;
    const float* v = valid + base;
    const float* c = collision + base;
    const size_t p = (size_t)j * W + i;
    const unsigned char m = (v[p] != c[p]) ? 1 : 0;
    unsigned char md = 0;
    for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
            const int jj = j + dj, ii = i + di;
            if (jj >= 0 && jj < H && ii >= 0 && ii < W) {
                const size_t q = (size_t)jj * W + ii;
                md |= (v[q] != c[q]) ? 1 : 0;
            }
        }
    const unsigned char hp = (unsigned char)(v[p] * (float)(md == m ? 1 : 0));  // (H * P).astype(uint8)
    mask[base + p] = (unsigned char)(1 - hp);
}

// ---- depth loaders' arithmetic (utils.get_depth / get_disparity + Convert.disparity_to_depth) ----------------
// float64 end to end as numpy / torch evaluate it; the float32 output is the float64 value rounded once.
template <typename SRC, typename OUT>
__global__ void __launch_bounds__(256) depth_from_png_kernel(const SRC* __restrict__ src, int kind, size_t n, OUT* __restrict__ out) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        double v = (double)src[p];
        double d;
        if (kind == OFD_SRC_RELDEPTH) {
            if (v > 240.0) v = 240.0;                      // utils.py:119
            d = __ddiv_rn(1.0, __dsub_rn(255.0, v));       // utils.py:120
        } else {
            const double disp = __ddiv_rn(__dmul_rn(v, 63.0), 255.0);  // utils.py:66
            // preprocess.py:258-261: `50 / (disparity + 0.005)` with an int numerator is Tensor.__rtruediv__, which torch
            // evaluates as reciprocal() * 50 - two roundings, reproduced here
            d = __dmul_rn(__ddiv_rn(1.0, __dadd_rn(disp, 0.005)), 50.0);
        }
        out[p] = (OUT)d;
    }
}

static unsigned blocks_for(size_t n, unsigned per_block, unsigned cap) {
    size_t g = (n + per_block - 1) / per_block;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ofd

using namespace ofd;

extern "C" {

int ofd_reproject_flow(const void* depth, int depth_dtype, const float* cam, float eps, int B, int H, int W,
                       float* flow, ofd_stream_t stream) {
    const char* fn = "ofd_reproject_flow";
    if (depth_dtype != OFD_F32 && depth_dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad depth dtype %d", fn, depth_dtype);
    if (B < 0 || H < 0 || W < 0 || B > 65535 || (H + 7) / 8 > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!depth || !cam || !flow) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    dim3 grid((W + 127) / 128, (H + 7) / 8, B), block(32, 8);
    if (depth_dtype == OFD_F32)
        reproject_flow_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)depth, (const Cam*)cam, eps, H, W, flow);
    else
        reproject_flow_kernel<double><<<grid, block, 0, (cudaStream_t)stream>>>((const double*)depth, (const Cam*)cam, eps, H, W, flow);
    return check_launch(fn);
}

int ofd_backproject(const void* depth, int depth_dtype, const float* invk3, int B, int H, int W, float* cam_points,
                    ofd_stream_t stream) {
    const char* fn = "ofd_backproject";
    if (depth_dtype != OFD_F32 && depth_dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad depth dtype %d", fn, depth_dtype);
    if (B < 0 || H < 0 || W < 0 || B > 65535 || (H + 7) / 8 > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!depth || !invk3 || !cam_points) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    dim3 grid((W + 31) / 32, (H + 7) / 8, B), block(32, 8);
    if (depth_dtype == OFD_F32)
        backproject_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)depth, invk3, H, W, cam_points);
    else
        backproject_kernel<double><<<grid, block, 0, (cudaStream_t)stream>>>((const double*)depth, invk3, H, W, cam_points);
    return check_launch(fn);
}

int ofd_project(const float* cam_points, const float* P, float eps, int B, int H, int W, float* pix_coords, float* z,
                ofd_stream_t stream) {
    const char* fn = "ofd_project";
    if (B < 0 || H < 0 || W < 0 || B > 65535 || (H + 7) / 8 > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!cam_points || !P || !pix_coords || !z) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    dim3 grid((W + 31) / 32, (H + 7) / 8, B), block(32, 8);
    project_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(cam_points, P, eps, H, W, pix_coords, z);
    return check_launch(fn);
}

int ofd_special_flow(int kind, const float* params_host, int H, int W, float* flow, float* back_flow,
                     ofd_stream_t stream) {
    const char* fn = "ofd_special_flow";
    if (kind < 5 || kind > 7) return fail(OFD_E_ARG, "%s: kind must be 5 (flip), 6 (rotate) or 7 (shear)", fn);
    if (H < 0 || W < 0 || (H + 7) / 8 > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (H == 0 || W == 0) return OFD_OK;
    if (!flow || !back_flow) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    SpecialParams sp = {};
    if (kind != 5) {
        if (!params_host) return fail(OFD_E_NULL, "%s: params_host is NULL", fn);
        // layout: cx, cy, M(4), Mrev(4); kind 7 ignores the centre
        sp.cx = params_host[0];
        sp.cy = params_host[1];
        for (int k = 0; k < 4; ++k) sp.m[k] = params_host[2 + k], sp.mrev[k] = params_host[6 + k];
        sp.use_center = (kind == 6);
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
    special_flow_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(kind, sp, H, W, flow, back_flow);
    return check_launch(fn);
}

int ofd_special_flow_batch(const int* kinds_host, const float* params_host, int B, int H, int W, float* flow,
                           float* back_flow, ofd_stream_t stream) {
    const char* fn = "ofd_special_flow_batch";
    if (B < 0 || H < 0 || W < 0 || (H + 7) / 8 > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!kinds_host || !params_host || !flow || !back_flow) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    for (int b = 0; b < B; ++b)
        if (kinds_host[b] < 5 || kinds_host[b] > 7)
            return fail(OFD_E_ARG, "%s: kinds[%d] = %d, must be 5 (flip), 6 (rotate) or 7 (shear)", fn, b, kinds_host[b]);
    const size_t hw = (size_t)H * W;
    for (int b0 = 0; b0 < B; b0 += SPECIAL_BATCH) {
        const int n = (B - b0) < SPECIAL_BATCH ? (B - b0) : SPECIAL_BATCH;
        SpecialBatch sb = {};
        for (int k = 0; k < n; ++k) {
            const float* q = params_host + (size_t)(b0 + k) * 10;
            sb.kind[k] = kinds_host[b0 + k];
            if (sb.kind[k] == 5) continue;
            sb.sp[k].cx = q[0];
            sb.sp[k].cy = q[1];
            for (int m = 0; m < 4; ++m) sb.sp[k].m[m] = q[2 + m], sb.sp[k].mrev[m] = q[6 + m];
            sb.sp[k].use_center = (sb.kind[k] == 6);
        }
        dim3 grid((W + 31) / 32, (H + 7) / 8, n), block(32, 8);
        special_flow_batch_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(sb, H, W, flow + (size_t)b0 * 2 * hw,
                                                                           back_flow + (size_t)b0 * 2 * hw);
        int rc = check_launch(fn);
        if (rc) return rc;
    }
    return OFD_OK;
}

int ofd_normalize_depth(const void* depth, int dtype, int B, int H, int W, void* out, void* scratch,
                        ofd_stream_t stream) {
    const char* fn = "ofd_normalize_depth";
    if (dtype != OFD_F32 && dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad dtype %d", fn, dtype);
    if (B < 0 || H < 0 || W < 0 || B > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!depth || !out || !scratch) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    if ((uintptr_t)scratch & 7) return fail(OFD_E_WORKSPACE, "%s: scratch must be 8-byte aligned (2*B uint64)", fn);
    const size_t hw = (size_t)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    u64* sc = (u64*)scratch;
    minmax_init_kernel<<<(B + 255) / 256, 256, 0, st>>>(sc, B);
    dim3 grid(blocks_for(hw, 256 * 8, 296), B);
    if (dtype == OFD_F32) {
        minmax_kernel<float><<<grid, 256, 0, st>>>((const float*)depth, hw, sc);
        normalize_map_kernel<float><<<grid, 256, 0, st>>>((const float*)depth, hw, sc, (float*)out);
    } else {
        minmax_kernel<double><<<grid, 256, 0, st>>>((const double*)depth, hw, sc);
        normalize_map_kernel<double><<<grid, 256, 0, st>>>((const double*)depth, hw, sc, (double*)out);
    }
    return check_launch(fn);
}

int ofd_normalize_depth_ragged(const void* depth, int dtype, int n_images, const size_t* count_host, const size_t* offset_host,
                               void* out, void* scratch, ofd_stream_t stream) {
    const char* fn = "ofd_normalize_depth_ragged";
    if (dtype != OFD_F32 && dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad dtype %d", fn, dtype);
    if (n_images < 0) return fail(OFD_E_SHAPE, "%s: negative image count", fn);
    if (n_images == 0) return OFD_OK;
    if (!depth || !out || !scratch || !count_host || !offset_host) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    if ((uintptr_t)scratch & 7) return fail(OFD_E_WORKSPACE, "%s: scratch must be 8-byte aligned (2*n_images uint64)", fn);
    cudaStream_t st = (cudaStream_t)stream;
    u64* sc = (u64*)scratch;
    minmax_init_kernel<<<(n_images + 255) / 256, 256, 0, st>>>(sc, n_images);
    for (int i0 = 0; i0 < n_images; i0 += RAGGED_MAX) {
        const int n = (n_images - i0) < RAGGED_MAX ? (n_images - i0) : RAGGED_MAX;
        RaggedTable tab = {};
        size_t big = 0;
        for (int k = 0; k < n; ++k) {
            tab.offset[k] = offset_host[i0 + k];
            tab.count[k] = count_host[i0 + k];
            big = count_host[i0 + k] > big ? count_host[i0 + k] : big;
        }
        if (big == 0) continue;
        dim3 grid(blocks_for(big, 256 * 8, 296), n);
        if (dtype == OFD_F32) {
            minmax_ragged_kernel<float><<<grid, 256, 0, st>>>((const float*)depth, tab, sc + 2 * i0);
            normalize_map_ragged_kernel<float><<<grid, 256, 0, st>>>((const float*)depth, tab, sc + 2 * i0, (float*)out);
        } else {
            minmax_ragged_kernel<double><<<grid, 256, 0, st>>>((const double*)depth, tab, sc + 2 * i0);
            normalize_map_ragged_kernel<double><<<grid, 256, 0, st>>>((const double*)depth, tab, sc + 2 * i0, (double*)out);
        }
        int rc = check_launch(fn);
        if (rc) return rc;
    }
    return OFD_OK;
}

int ofd_depth_from_png(const void* src, int src_bits, int kind, size_t n, void* depth, int depth_dtype, ofd_stream_t stream) {
    const char* fn = "ofd_depth_from_png";
    if (src_bits != 8 && src_bits != 16) return fail(OFD_E_DTYPE, "%s: src_bits must be 8 or 16", fn);
    if (kind != OFD_SRC_RELDEPTH && kind != OFD_SRC_DISPARITY) return fail(OFD_E_ARG, "%s: bad kind %d", fn, kind);
    if (depth_dtype != OFD_F32 && depth_dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad depth dtype %d", fn, depth_dtype);
    if (n == 0) return OFD_OK;
    if (!src || !depth) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = blocks_for(n, 256 * 4, 148 * 8);
    if (src_bits == 8) {
        if (depth_dtype == OFD_F64) depth_from_png_kernel<uint8_t, double><<<g, 256, 0, st>>>((const uint8_t*)src, kind, n, (double*)depth);
        else depth_from_png_kernel<uint8_t, float><<<g, 256, 0, st>>>((const uint8_t*)src, kind, n, (float*)depth);
    } else {
        if (depth_dtype == OFD_F64) depth_from_png_kernel<uint16_t, double><<<g, 256, 0, st>>>((const uint16_t*)src, kind, n, (double*)depth);
        else depth_from_png_kernel<uint16_t, float><<<g, 256, 0, st>>>((const uint16_t*)src, kind, n, (float*)depth);
    }
    return check_launch(fn);
}

int ofd_fix_warped_depth(float* depth, size_t n, ofd_stream_t stream) {
    if (n == 0) return OFD_OK;
    if (!depth) return fail(OFD_E_NULL, "ofd_fix_warped_depth: NULL pointer");
    fix_depth_kernel<<<blocks_for(n, 256 * 4, 148 * 8), 256, 0, (cudaStream_t)stream>>>(depth, n);
    return check_launch("ofd_fix_warped_depth");
}

int ofd_inpaint_mask(const float* valid, const float* collision, int B, int H, int W, uint8_t* mask, ofd_stream_t stream) {
    const char* fn = "ofd_inpaint_mask";
    if (B < 0 || H < 0 || W < 0 || B > 65535 || (H + 7) / 8 > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!valid || !collision || !mask) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    dim3 grid((W + 31) / 32, (H + 7) / 8, B), block(32, 8);
    inpaint_mask_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(valid, collision, H, W, mask);
    return check_launch(fn);
}

}  // extern "C"
