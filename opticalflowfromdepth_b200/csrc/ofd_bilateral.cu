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
// A 32x8 output tile stages depth and the discontinuity flag for tile + halo in shared memory with clamp
// addressing (the replicated ring makes every window tap read row clamp(r,1,H-2), col clamp(c,1,W-2)); the
// selection is a value-rank count over the window, only run by pixels whose window holds a discontinuity.
#include "ofd_common.cuh"

namespace ofd {

constexpr int BT_W = 32, BT_H = 8, MAX_WIN = 15;

template <typename DT>
__global__ void __launch_bounds__(BT_W* BT_H) bilateral_iter_kernel(const DT* __restrict__ din, const DT* __restrict__ dorig,
                                                                   int H, int W, int win, DT thr, DT* __restrict__ dout) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = win / 2;
    const int TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    DT* sdep = reinterpret_cast<DT*>(smem_raw);
    unsigned char* sdisc = reinterpret_cast<unsigned char*>(sdep + TW * TH);
    __shared__ int krank[MAX_WIN * MAX_WIN + 1];
    const int tid = threadIdx.y * BT_W + threadIdx.x;
    const int r0 = blockIdx.y * BT_H - m, c0 = blockIdx.x * BT_W - m;

    // rank table k(n), numpy float32 semantics (bilateral_filter.py:194-197)
    for (int n = 1 + tid; n <= win * win; n += BT_W * BT_H) {
        const float w = __fdiv_rn(1.0f, (float)n);
        float cum = 0.0f;
        int k = 0;
        for (int q = 0; q < n; ++q) {
            cum = __fadd_rn(cum, w);
            k += (cum <= 0.5f);
        }
        krank[n] = k;
    }
    // stage depth + discontinuity for tile + halo, at ring-replicated coordinates
    for (int e = tid; e < TW * TH; e += BT_W * BT_H) {
        const int tr = e / TW, tc = e - tr * TW;
        int r = r0 + tr, c = c0 + tc;
        r = r < 1 ? 1 : (r > H - 2 ? H - 2 : r);
        c = c < 1 ? 1 : (c > W - 2 ? W - 2 : c);
        const size_t p = (size_t)r * W + c;
        const DT d = din[p];
        const DT inv = (DT)1.0 / d;
        // (r,c) is interior by construction: all four neighbours exist
        const DT iu = (DT)1.0 / din[p - W], ib = (DT)1.0 / din[p + W];
        const DT il = (DT)1.0 / din[p - 1], ir = (DT)1.0 / din[p + 1];
        bool disc = (fabs(inv - iu) > thr) || (fabs(inv - ib) > thr) || (fabs(inv - il) > thr) || (fabs(inv - ir) > thr);
        disc = disc || (dorig[p] == (DT)0);
        sdep[e] = d;
        sdisc[e] = disc ? 1 : 0;
    }
    __syncthreads();
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
    dout[(size_t)r * W + c] = result;
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
    const int m = window / 2;
    const int TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    const size_t esz = dtype == OFD_F32 ? 4 : 8;
    const size_t smem = (size_t)TW * TH * (esz + 1) + 16;
    dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H), block(BT_W, BT_H);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OFD_F32)
        bilateral_iter_kernel<float><<<grid, block, smem, st>>>((const float*)depth_in, (const float*)depth_orig, H, W, window,
                                                                (float)threshold, (float*)depth_out);
    else
        bilateral_iter_kernel<double><<<grid, block, smem, st>>>((const double*)depth_in, (const double*)depth_orig, H, W, window,
                                                                 threshold, (double*)depth_out);
    return check_launch(fn);
}
