// ofd_bilateral.cu — one iteration of the reference's "sparse bilateral filtering", which on the path the
// reference takes (discontinuity_map given, mask None) is a depth-edge-gated MEDIAN (bilateral_filter.py:33-58):
//
//   disc(r,c) = any 4-neighbour |1/d - 1/d'| > thr, interior pixels only            bilateral_filter.py:63-116
//   disc[depth_orig == 0] = 1                                                        :46
//   depth, disc: border ring replaced by the edge-replicated interior, then edge-padded by window/2   :141-147
//   for every pixel whose window holds a discontinuity: out = rank-k(n) smallest depth among the n window
//   pixels with disc == 0 (centre value when n == 0); all other pixels keep the (ring-replicated) depth   :167-198
//   k(n) = #{ m in 1..n : float32 running sum of m copies of float32(1/n) <= 0.5 }   :194-197 (cumsum + digitize)
//
// Kernel plan (one CTA = a 32x8 output tile, all staging in shared memory, 4-12 B/px of HBM traffic):
//   1. the raw depth tile with a halo of window/2 + 2 is loaded once (zero outside the image) and 1/d is formed once
//      per cell - not once per tap;
//   2. the discontinuity flag of every interior cell comes from its four shared-memory neighbours;
//   3. a replication pass applies the reference's border rule (every tap reads row clamp(r,1,H-2), column
//      clamp(c,1,W-2)) so the window taps are plain offsets;
//   4. pixels whose window holds a discontinuity pull their window into REGISTERS (discontinuity taps -> +inf) and
//      sort it with a fully unrolled bitonic network (window 3/5/7: 16/32/64 keys, 80/240/672 compare-exchanges -
//      against ~2400 x 2 shared-memory compares of a rank count); other window sizes use the rank count.
// A TMA 2-D tiled load was considered for step 1 (the north star suggests it); a 40x16 float tile is 2.5 loads per
// thread, so plain coalesced loads are as fast and need no tensor map per call - the time goes into step 4.
#include "ofd_common.cuh"

namespace ofd {

constexpr int BT_W = 32, BT_H = 8, MAX_WIN = 15;

template <typename DT>
__device__ __forceinline__ DT pos_inf();
template <>
__device__ __forceinline__ float pos_inf<float>() {
    return __int_as_float(0x7f800000);
}
template <>
__device__ __forceinline__ double pos_inf<double>() {
    return __longlong_as_double(0x7ff0000000000000ll);
}

// fully unrolled bitonic sort (ascending) of N = 2^k register values
template <typename DT, int N>
__device__ __forceinline__ void bitonic_sort(DT (&v)[N]) {
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ stride;
                if (l > i) {
                    const bool up = ((i & size) == 0);
                    const DT a = v[i], b = v[l];
                    const DT lo = a < b ? a : b, hi = a < b ? b : a;
                    v[i] = up ? lo : hi;
                    v[l] = up ? hi : lo;
                }
            }
        }
    }
}

// WS > 0: compile-time window (register sort); WS == 0: run-time window (rank count in shared memory)
template <typename DT, int WS>
__global__ void __launch_bounds__(BT_W* BT_H) bilateral_iter_kernel(const DT* __restrict__ din, const DT* __restrict__ dorig,
                                                                   int H, int W, int win_rt, DT thr, DT* __restrict__ dout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int win = WS > 0 ? WS : win_rt;
    const int m = win / 2;
    const int RW = BT_W + 2 * m + 4, RH = BT_H + 2 * m + 4;  // raw tile: window halo + a ring of 2 (neighbour differences + clamping)
    const int TW = BT_W + 2 * m, TH = BT_H + 2 * m;          // window tile
    DT* sraw = reinterpret_cast<DT*>(smem_raw);               // raw depth
    DT* sinv = sraw + RW * RH;                                // 1 / depth, formed once per cell
    DT* sdep = sinv + RW * RH;                                // replicated depth at window coordinates
    unsigned char* sflag = reinterpret_cast<unsigned char*>(sdep + TW * TH);  // raw-tile discontinuity flags
    unsigned char* sdisc = sflag + RW * RH;                                   // replicated flags at window coordinates
    __shared__ int krank[MAX_WIN * MAX_WIN + 1];
    const int tid = threadIdx.y * BT_W + threadIdx.x;
    const int nthr = BT_W * BT_H;
    const int r0 = blockIdx.y * BT_H - m - 2, c0 = blockIdx.x * BT_W - m - 2;  // raw coordinate of raw-tile cell (0,0)

    // rank table k(n), numpy float32 semantics (bilateral_filter.py:194-197)
    for (int n = 1 + tid; n <= win * win; n += nthr) {
        const float w = __fdiv_rn(1.0f, (float)n);
        float cum = 0.0f;
        int k = 0;
        for (int q = 0; q < n; ++q) {
            cum = __fadd_rn(cum, w);
            k += (cum <= 0.5f);
        }
        krank[n] = k;
    }
    // 1. raw tile: depth, 1/depth, depth_orig == 0 (zero outside the image; those cells are never consumed)
    for (int e = tid; e < RW * RH; e += nthr) {
        const int tr = e / RW, tc = e - tr * RW;
        const int r = r0 + tr, c = c0 + tc;
        DT d = (DT)0;
        unsigned char z = 0;
        if (r >= 0 && r < H && c >= 0 && c < W) {
            const size_t p = (size_t)r * W + c;
            d = din[p];
            z = (dorig[p] == (DT)0) ? 2 : 0;  // bit 1: forced discontinuity (:46)
        }
        sraw[e] = d;
        sinv[e] = (DT)1.0 / d;
        sflag[e] = z;
    }
    __syncthreads();
    // 2. discontinuity of interior image pixels from their four shared-memory neighbours (:63-116)
    for (int e = tid; e < RW * RH; e += nthr) {
        const int tr = e / RW, tc = e - tr * RW;
        const int r = r0 + tr, c = c0 + tc;
        if (tr > 0 && tr < RH - 1 && tc > 0 && tc < RW - 1 && r >= 1 && r <= H - 2 && c >= 1 && c <= W - 2) {
            const DT inv = sinv[e];
            const bool disc = (fabs(inv - sinv[e - RW]) > thr) || (fabs(inv - sinv[e + RW]) > thr) ||
                              (fabs(inv - sinv[e - 1]) > thr) || (fabs(inv - sinv[e + 1]) > thr);
            if (disc) sflag[e] |= 1;
        }
    }
    __syncthreads();
    // 3. replication pass: window cell (tr,tc) <- raw cell of the clamped coordinate (border ring rule, :141-147)
    for (int e = tid; e < TW * TH; e += nthr) {
        const int tr = e / TW, tc = e - tr * TW;
        int r = r0 + 2 + tr, c = c0 + 2 + tc;
        r = r < 1 ? 1 : (r > H - 2 ? H - 2 : r);
        c = c < 1 ? 1 : (c > W - 2 ? W - 2 : c);
        const int src = (r - r0) * RW + (c - c0);
        sdep[e] = sraw[src];
        sdisc[e] = sflag[src] ? 1 : 0;
    }
    __syncthreads();
    // 4. gated median
    const int r = blockIdx.y * BT_H + threadIdx.y, c = blockIdx.x * BT_W + threadIdx.x;
    if (r >= H || c >= W) return;
    const int base = threadIdx.y * TW + threadIdx.x;  // top-left tap of this pixel's window
    const DT centre = sdep[base + m * TW + m];
    int n_disc = 0;
    for (int dr = 0; dr < win; ++dr)
        for (int dc = 0; dc < win; ++dc) n_disc += sdisc[base + dr * TW + dc];
    DT result = centre;
    const int n = win * win - n_disc;
    if (n_disc > 0 && n > 0) {
        const int k = krank[n];
        if (WS > 0) {
            constexpr int NN = WS * WS <= 16 ? 16 : (WS * WS <= 32 ? 32 : 64);
            DT v[NN];
#pragma unroll
            for (int q = 0; q < NN; ++q) {
                DT x = pos_inf<DT>();
                if (q < WS * WS) {
                    const int e = base + (q / WS) * TW + (q % WS);
                    if (!sdisc[e]) x = sdep[e];
                }
                v[q] = x;
            }
            bitonic_sort<DT, NN>(v);
            result = v[0];
#pragma unroll
            for (int q = 1; q < WS * WS; ++q) result = (q == k) ? v[q] : result;
        } else {
            // value with  #{f < v} <= k < #{f <= v}  among the n non-discontinuity taps
            for (int er = 0; er < win; ++er) {
                for (int ec = 0; ec < win; ++ec) {
                    const int e = base + er * TW + ec;
                    if (sdisc[e]) continue;
                    const DT v = sdep[e];
                    int lt = 0, le = 0;
                    for (int dr = 0; dr < win; ++dr)
                        for (int dc = 0; dc < win; ++dc) {
                            const int f = base + dr * TW + dc;
                            if (!sdisc[f]) {
                                const DT u = sdep[f];
                                lt += (u < v);
                                le += (u <= v);
                            }
                        }
                    if (lt <= k && k < le) {
                        result = v;
                        er = win;  // done
                        break;
                    }
                }
            }
        }
    }
    dout[(size_t)r * W + c] = result;
}

template <typename DT, int WS>
static void launch_bilateral(const DT* din, const DT* dorig, int H, int W, int window, DT thr, DT* dout, cudaStream_t st) {
    const int m = window / 2;
    const int RW = BT_W + 2 * m + 4, RH = BT_H + 2 * m + 4, TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    const size_t smem = (size_t)RW * RH * (2 * sizeof(DT) + 1) + (size_t)TW * TH * (sizeof(DT) + 1) + 32;
    dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H), block(BT_W, BT_H);
    bilateral_iter_kernel<DT, WS><<<grid, block, smem, st>>>(din, dorig, H, W, window, thr, dout);
}

template <typename DT>
static void dispatch_bilateral(const DT* din, const DT* dorig, int H, int W, int window, DT thr, DT* dout, cudaStream_t st) {
    switch (window) {
        case 3: launch_bilateral<DT, 3>(din, dorig, H, W, window, thr, dout, st); break;
        case 5: launch_bilateral<DT, 5>(din, dorig, H, W, window, thr, dout, st); break;
        case 7: launch_bilateral<DT, 7>(din, dorig, H, W, window, thr, dout, st); break;
        default: launch_bilateral<DT, 0>(din, dorig, H, W, window, thr, dout, st); break;
    }
}

}  // namespace ofd

using namespace ofd;

extern "C" int ofd_bilateral_iter(const void* depth_in, const void* depth_orig, int dtype, int H, int W, int window,
                                  double threshold, void* depth_out, ofd_stream_t stream) {
    const char* fn = "ofd_bilateral_iter";
    if (dtype != OFD_F32 && dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad dtype %d", fn, dtype);
    if (H < 3 || W < 3) return fail(OFD_E_SHAPE, "%s: needs H >= 3 and W >= 3 (the reference pads an empty interior otherwise)", fn);
    if ((H + BT_H - 1) / BT_H > 65535) return fail(OFD_E_SHAPE, "%s: H too large", fn);
    if (window < 1 || window > MAX_WIN || (window & 1) == 0)
        return fail(OFD_E_ARG, "%s: window must be odd and in [1,%d]", fn, MAX_WIN);
    if (!depth_in || !depth_orig || !depth_out) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    if (depth_in == depth_out) return fail(OFD_E_ARG, "%s: in-place filtering is not supported", fn);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OFD_F32)
        dispatch_bilateral<float>((const float*)depth_in, (const float*)depth_orig, H, W, window, (float)threshold, (float*)depth_out, st);
    else
        dispatch_bilateral<double>((const double*)depth_in, (const double*)depth_orig, H, W, window, threshold, (double*)depth_out, st);
    return check_launch(fn);
}
