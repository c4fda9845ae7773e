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
// Kernel plan (one CTA of 256 threads = a 32x32 output tile, all staging in shared memory, 4-12 B/px of HBM traffic):
//   1. the raw depth tile with a halo of window/2 + 2 is loaded once (zero outside the image; every global load of a thread
//      is issued before the first use) and 1/d is formed once per cell - not once per tap;
//   2. the discontinuity flag of every interior cell comes from its four shared-memory neighbours;
//   3. tiles on the image border run a replication pass that applies the reference's border rule (every tap reads row
//      clamp(r,1,H-2), column clamp(c,1,W-2)); interior tiles read the raw tile in place;
//   4. one 64-bit discontinuity mask per window row (ballots); every thread counts the windows of its 4 consecutive rows
//      with sliding sums; pixels whose window holds a discontinuity are COMPACTED into a per-tile list (so the selection
//      runs in full warps), pull their window into REGISTERS (discontinuity taps -> +inf) and sort it with a fully unrolled
//      odd-even merge-sort network over exactly 9/25/49 keys (28/140/394 compare-exchanges of 2 FMNMX each; the half of the
//      outputs that can never be selected is dead code); other window sizes use a rank count in shared memory.
// Variants: a ragged batch (one launch for images of different sizes) and the reference's binary-mask path.
// The kernel is instruction-bound, not HBM-bound.  The north star's "TMA-loaded halo tiles" were built and MEASURED
// (OFD_BIL_TMA=1: both raw tiles by cp.async.bulk.tensor.2d, SASS UTMALDG, 1/d and forced flags formed in a pass over shared memory;
// bit-identical): 5 iterations on one dense frame take 0.124 ms against 0.117 ms at 1080p (6 % slower) and 0.345 against 0.368 ms at
// 2160x3840 (6 % faster) - profiles/r2/tune_bilateral_tma.txt.  The per-thread loads it replaces are 7-8 coalesced LDGs per thread whose
// address arithmetic is a few instructions per cell; the TMA variant trades them for a second pass over shared memory, a wider tile (the box
// must start on a 16-byte boundary of the row: 48 instead of 42 columns at window 7) and one tensor map per image (ragged batches would need
// one per frame).  A wash: the SIMT staging stays the default, the TMA path stays behind OFD_BIL_TMA for dense float32 frames.
#include <cuda.h>  // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint, no -lcuda)

#include <cstdlib>

#include "ofd_common.cuh"

namespace ofd {

// ---- TMA 2-D tiled load of the raw tile (EXPERIMENT, OFD_BIL_TMA=1; VERDICT r1 next #7b) -----------------------------------------------
__device__ __forceinline__ uint32_t bil_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bil_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bil_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bil_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bil_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bil_mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}" ::"r"(bil_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one box of the tensor map at element coordinates (x, y) - out-of-image elements arrive as zeros
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     bil_smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(bil_smem_u32(bar))
                 : "memory");
}

// Tile = 32 x (8 * BIL_RPT) outputs for a block of 32 x 8 threads: every thread owns BIL_RPT output rows.  A taller tile
// amortises the halo (window 7: 2.95 staged cells per output at 32x8, 2.13 at 32x16, 1.72 at 32x32).
#ifndef OFD_BIL_RPT
#define OFD_BIL_RPT 4
#endif
// float kernels: 6 CTAs per SM (<= 42 registers) hide the tile-load latency better than 3 CTAs at 72 registers (cfg2 step
// 1.19 -> 1.06 ms); the double kernels keep their registers (the 49-key network alone holds 98 of them).
#ifndef OFD_BIL_MINB
#define OFD_BIL_MINB 6
#endif
constexpr int BIL_RPT = OFD_BIL_RPT;
constexpr int BT_W = 32, BT_TY = 8, BT_H = BT_TY * BIL_RPT, MAX_WIN = 15;

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

// Batcher's odd-even merge sort as a compile-time comparator list for ANY n (not only powers of two): 28 / 140 / 394
// compare-exchanges for 9 / 25 / 49 keys (a bitonic network over the padded 16 / 32 / 64 keys needs 80 / 240 / 672).
// Only ranks 0 .. (n-1)/2 are ever selected, so the compiler also drops the comparators that feed higher outputs only
// (25 / 124 / 352 remain).  Each compare-exchange is two FMNMX (min / max) instead of FSETP + 2 FSEL.
struct CePair {
    int a, b;
};
// the idx-th comparator of the network over n keys ({-1,-1} past the end); evaluated at compile time only
constexpr CePair oems_pair(int n, int idx) {
    int count = 0;
    for (int p = 1; p < n; p *= 2)
        for (int k = p; k >= 1; k /= 2)
            for (int j = k % p; j <= n - 1 - k; j += 2 * k) {
                const int imax = (k - 1 < n - j - k - 1) ? k - 1 : n - j - k - 1;
                for (int i = 0; i <= imax; ++i)
                    if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        if (count == idx) return CePair{i + j, i + j + k};
                        ++count;
                    }
            }
    return CePair{-1, -1};
}
constexpr int oems_count(int n) {
    int count = 0;
    while (oems_pair(n, count).a >= 0) ++count;
    return count;
}

template <typename DT>
__device__ __forceinline__ void compare_exchange(DT& x, DT& y);
template <>
__device__ __forceinline__ void compare_exchange<float>(float& x, float& y) {
    const float lo = fminf(x, y), hi = fmaxf(x, y);
    x = lo, y = hi;
}
template <>
__device__ __forceinline__ void compare_exchange<double>(double& x, double& y) {
    const double lo = fmin(x, y), hi = fmax(x, y);
    x = lo, y = hi;
}

// comparators [LO, HI) of the network, unrolled by halving (keeps the template recursion depth at log2 of the count)
template <typename DT, int N, int LO, int HI>
__device__ __forceinline__ void oems_range(DT (&v)[N]) {
    if constexpr (HI - LO == 1) {
        constexpr CePair ce = oems_pair(N, LO);
        compare_exchange<DT>(v[ce.a], v[ce.b]);
    } else if constexpr (HI - LO > 1) {
        oems_range<DT, N, LO, (LO + HI) / 2>(v);
        oems_range<DT, N, (LO + HI) / 2, HI>(v);
    }
}

// ascending sort of N register values (N compile-time, any value)
template <typename DT, int N>
__device__ __forceinline__ void network_sort(DT (&v)[N]) {
    oems_range<DT, N, 0, oems_count(N)>(v);
}

// rank table k(n) = #{ m in 1..n : float32 running sum of m copies of float32(1/n) <= 0.5 } (numpy cumsum + digitize,
// bilateral_filter.py:194-197), n <= 15 x 15.  Filled on the host once per call and passed in the parameter block.
struct KRank {
    unsigned char k[MAX_WIN * MAX_WIN + 3];
};

// coef_f64: the coefficients are float64 (mask path with a float64 / integer mask, bilateral_filter.py:182), else float32
static KRank make_krank(int window, bool coef_f64 = false) {
    KRank kr = {};
    for (int n = 1; n <= window * window; ++n) {
        int k = 0;
        if (coef_f64) {
            const double w = 1.0 / (double)n;
            volatile double cum = 0.0;
            for (int q = 0; q < n; ++q) {
                cum = cum + w;
                k += (cum <= 0.5);
            }
        } else {
            const float w = 1.0f / (float)n;
            volatile float cum = 0.0f;  // volatile: every partial sum is rounded to float32
            for (int q = 0; q < n; ++q) {
                cum = cum + w;
                k += (cum <= 0.5f);
            }
        }
        kr.k[n] = (unsigned char)k;
    }
    return kr;
}

// WS > 0: compile-time window (register sort); WS == 0: run-time window (rank count in shared memory)
// MASK: the reference's binary-mask path (bilateral_filter.py:48-49,72-80,161,169-170,180-182; WS > 0 only): a neighbour difference
// counts only between two unmasked pixels, masked pixels are never discontinuities and keep their depth, masked taps and taps
// outside the image (the mask is zero-padded, not ring-replicated) are left out of the median.  Flag byte: bit 0 discontinuity,
// bit 1 depth_orig == 0, bit 2 unmasked.
template <typename DT, int WS, bool MASK = false, bool TMA = false>
__device__ __forceinline__ void bilateral_tile(const DT* __restrict__ din, const DT* __restrict__ dorig, int H, int W,
                                               int win_rt, DT thr, DT* __restrict__ dout, int tile_x, int tile_y,
                                               const KRank& kr, const unsigned char* __restrict__ mask = nullptr,
                                               const CUtensorMap* tm_in = nullptr, const CUtensorMap* tm_orig = nullptr) {
    static_assert(!MASK || WS > 0, "the mask path is built for the compile-time windows");
    static_assert(!TMA || (WS > 0 && !MASK && sizeof(DT) == 4), "the TMA experiment covers the float32 compile-time windows");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int win = WS > 0 ? WS : win_rt;
    const int m = win / 2;
    // raw tile: window halo + a ring of 2 (neighbour differences + clamping).  A TMA box must START on a 16-byte boundary of the tensor row
    // and span a multiple of 16 bytes (an unaligned start coordinate raises an illegal-instruction fault: tools/probes/tma_probe.cu), so the
    // TMA variant widens the tile by LP columns on the left and rounds the row up to a multiple of 4 floats; CX = first window-tile column
    const int LP = TMA ? (4 - ((m + 2) & 3)) & 3 : 0, CX = 2 + LP;
    const int RW = TMA ? ((BT_W + 2 * m + 4 + LP + 3) & ~3) : BT_W + 2 * m + 4, RH = BT_H + 2 * m + 4;
    const int TW = BT_W + 2 * m, TH = BT_H + 2 * m;          // window tile
    const int CELLS_P = TMA ? ((RW * RH + 31) & ~31) : RW * RH;  // TMA destinations start on 128-byte boundaries
    DT* sraw = reinterpret_cast<DT*>(smem_raw);               // raw depth
    DT* sinv = sraw + CELLS_P;                                // 1 / depth, formed once per cell
    DT* sdep = sinv + CELLS_P;                                // replicated depth at window coordinates (border tiles only)
    unsigned char* sflag = reinterpret_cast<unsigned char*>(sdep + TW * TH);  // raw-tile discontinuity flags
    unsigned char* sdisc = sflag + RW * RH;                                   // replicated flags (border tiles only)
    __shared__ unsigned char s_k[MAX_WIN * MAX_WIN + 3];
    __shared__ unsigned long long s_rowmask[BT_H + MAX_WIN - 1];  // discontinuity bits of every window-tile row (TW <= 46 columns)
    __shared__ unsigned long long s_rowexcl[MASK ? BT_H + MAX_WIN - 1 : 1];  // MASK: taps left out of the median (discontinuity | masked)
    __shared__ unsigned short s_list[BT_W * BT_H];                // tile pixels that need the median (row * BT_W + column)
    __shared__ int s_count;
    const int tid = threadIdx.y * BT_W + threadIdx.x;
    constexpr int nthr = BT_W * BT_TY;
    if (tid == 0) s_count = 0;
    if (tid <= MAX_WIN * MAX_WIN) s_k[tid] = kr.k[tid];
    const int r0 = tile_y * BT_H - m - 2, c0 = tile_x * BT_W - m - 2 - LP;  // raw coordinate of raw-tile cell (0,0)

    // 1. raw tile: depth, 1/depth, depth_orig == 0 (zero outside the image; those cells are never consumed).  For a
    //    compile-time window all global loads of a thread are issued before the first use.
    if constexpr (TMA) {
        // EXPERIMENT: both raw tiles (depth, depth_orig) arrive by ONE tensor-map copy each (cp.async.bulk.tensor.2d, SASS UTMALDG; cells
        // outside the image are zero-filled by the TMA unit); the threads then form 1/d and the forced flags in a pass over shared memory.
        __shared__ __align__(8) uint64_t s_bar;
        if (tid == 0) bil_mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            bil_mbar_expect_tx(&s_bar, 2u * (unsigned)(RW * RH) * (unsigned)sizeof(DT));
            tma_load_2d(sraw, tm_in, c0, r0, &s_bar);
            tma_load_2d(sinv, tm_orig, c0, r0, &s_bar);
        }
        bil_mbar_wait(&s_bar, 0);
        for (int e = tid; e < RW * RH; e += nthr) {
            const int tr = e / RW, tc = e - tr * RW;
            const int r = r0 + tr, c = c0 + tc;
            const bool inb = r >= 0 && r < H && c >= 0 && c < W;
            const DT ov = sinv[e];
            sflag[e] = (inb && ov == (DT)0) ? 2 : 0;  // out-of-image cells arrive as zeros: they are not depth_orig == 0 pixels
            sinv[e] = (DT)1.0 / sraw[e];
        }
    } else if constexpr (WS > 0) {
        constexpr int CELLS = (BT_W + 2 * (WS / 2) + 4) * (BT_H + 2 * (WS / 2) + 4);
        constexpr int NL = (CELLS + nthr - 1) / nthr;
        DT dv[NL], ov[NL];
        unsigned char mv[MASK ? NL : 1];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const int e = tid + k * nthr;
            const int tr = e / RW, tc = e - tr * RW;
            const int r = r0 + tr, c = c0 + tc;
            dv[k] = (DT)0, ov[k] = (DT)1;
            if (MASK) mv[k] = 0;
            if (e < CELLS && r >= 0 && r < H && c >= 0 && c < W) {
                const size_t p = (size_t)r * W + c;
                dv[k] = din[p];
                ov[k] = dorig[p];
                if (MASK) mv[k] = mask[p];
            }
        }
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const int e = tid + k * nthr;
            if (e < CELLS) {
                sraw[e] = dv[k];
                sinv[e] = (DT)1.0 / dv[k];
                unsigned char f = (ov[k] == (DT)0) ? 2 : 0;  // bit 1: forced discontinuity (:46)
                if (MASK) f |= mv[k] ? 4 : 0;                // bit 2: unmasked (zero outside the image: the zero padding of :161)
                sflag[e] = f;
            }
        }
    } else {
        for (int e = tid; e < RW * RH; e += nthr) {
            const int tr = e / RW, tc = e - tr * RW;
            const int r = r0 + tr, c = c0 + tc;
            DT d = (DT)0;
            unsigned char z = 0;
            if (r >= 0 && r < H && c >= 0 && c < W) {
                const size_t p = (size_t)r * W + c;
                d = din[p];
                z = (dorig[p] == (DT)0) ? 2 : 0;
            }
            sraw[e] = d;
            sinv[e] = (DT)1.0 / d;
            sflag[e] = z;
        }
    }
    __syncthreads();
    // 2. discontinuity of interior image pixels from their four shared-memory neighbours (:63-116)
    int any_flag = 0;  // this thread saw a discontinuity (computed or forced) somewhere in the raw tile
    for (int e = tid; e < RW * RH; e += nthr) {
        const int tr = e / RW, tc = e - tr * RW;
        const int r = r0 + tr, c = c0 + tc;
        const bool tested = tr > 0 && tr < RH - 1 && tc > 0 && tc < RW - 1 && r >= 1 && r <= H - 2 && c >= 1 && c <= W - 2;
        if constexpr (!MASK) {
            any_flag |= sflag[e];  // forced discontinuities (depth_orig == 0) count as well
            if (tested) {
                // branch-free: the four comparisons are evaluated and OR-ed (| not ||), one predicated store
                const DT inv = sinv[e];
                const bool disc = (fabs(inv - sinv[e - RW]) > thr) | (fabs(inv - sinv[e + RW]) > thr) |
                                  (fabs(inv - sinv[e - 1]) > thr) | (fabs(inv - sinv[e + 1]) > thr);
                if (disc) sflag[e] |= 1;
                any_flag |= disc;
            }
        } else {
            // a difference counts only between two unmasked pixels (:72-80); then depth_orig == 0 forces (:46) and the mask clears
            // (:48-49).  Bit 2 of a neighbour is stable while other threads rewrite bits 0-1 of their own cells.
            const unsigned char fe = sflag[e];
            bool disc = false;
            if (tested) {
                const DT inv = sinv[e];
                disc = ((fabs(inv - sinv[e - RW]) > thr) & ((sflag[e - RW] & 4) != 0)) | ((fabs(inv - sinv[e + RW]) > thr) & ((sflag[e + RW] & 4) != 0)) |
                       ((fabs(inv - sinv[e - 1]) > thr) & ((sflag[e - 1] & 4) != 0)) | ((fabs(inv - sinv[e + 1]) > thr) & ((sflag[e + 1] & 4) != 0));
            }
            const unsigned char fin = (fe & 4) | (((fe & 4) && (disc || (fe & 2))) ? 1 : 0);  // bit 0 = final discontinuity, bit 1 dropped
            sflag[e] = fin;
            any_flag |= fin & 1;
        }
    }
    // 2b. block-uniform early-out: a tile without a single discontinuity in its raw tile (a superset of every window of the tile,
    //     ring replication included: clamped taps stay inside the raw tile) keeps its (ring-replicated) depth - steps 3-4 (replication
    //     pass, row masks, window counts, compaction, selection) are skipped.  Smooth regions are most of a depth map.
    if (!__syncthreads_or(any_flag)) {
        const int c = tile_x * BT_W + threadIdx.x;
#pragma unroll
        for (int rr = 0; rr < BIL_RPT; ++rr) {
            const int r = tile_y * BT_H + threadIdx.y * BIL_RPT + rr;
            if (r < H && c < W) {
                const int rc = r < 1 ? 1 : (r > H - 2 ? H - 2 : r), cc = c < 1 ? 1 : (c > W - 2 ? W - 2 : c);
                dout[(size_t)r * W + c] = sraw[(rc - r0) * RW + (cc - c0)];
            }
        }
        return;
    }
    // 3. the reference's border rule (:141-147): every tap reads row clamp(r,1,H-2), column clamp(c,1,W-2).  A tile whose
    //    window tile lies inside [1,H-2] x [1,W-2] needs no clamping: its taps read the raw tile in place.  Only tiles
    //    on the image border run the replication pass.
    const bool interior = (r0 + 2 >= 1) && (r0 + 2 + TH - 1 <= H - 2) && (c0 + CX >= 1) && (c0 + CX + TW - 1 <= W - 2);
    const DT* wdep = sraw + 2 * RW + CX;             // window-tile cell (tr,tc) -> wdep[tr * wstride + tc]
    const unsigned char* wdisc = sflag + 2 * RW + CX;  // non-zero = discontinuity
    int wstride = RW;
    if (!interior) {  // block-uniform
        for (int e = tid; e < TW * TH; e += nthr) {
            const int tr = e / TW, tc = e - tr * TW;
            int r = r0 + 2 + tr, c = c0 + CX + tc;
            r = r < 1 ? 1 : (r > H - 2 ? H - 2 : r);
            c = c < 1 ? 1 : (c > W - 2 ? W - 2 : c);
            const int src = (r - r0) * RW + (c - c0);
            sdep[e] = sraw[src];
            sdisc[e] = MASK ? (sflag[src] & 1) : sflag[src];
        }
        wdep = sdep, wdisc = sdisc, wstride = TW;
        __syncthreads();
    }
    // 4a. one 64-bit discontinuity mask per window-tile row (ballots), so a pixel tests its whole window with `win` loads
    {
        const int lane = tid & 31, warp = tid >> 5;
        constexpr unsigned char kDiscBits = MASK ? 1 : 0xFF;  // MASK: bit 2 of a raw flag byte is the mask, not a discontinuity
        for (int tr = warp; tr < TH; tr += nthr / 32) {
            const bool d_lo = (wdisc[tr * wstride + lane] & kDiscBits) != 0;
            const bool d_hi = (32 + lane < TW) && (wdisc[tr * wstride + 32 + lane] & kDiscBits) != 0;
            const unsigned lo = __ballot_sync(0xFFFFFFFFu, d_lo);
            const unsigned hi = __ballot_sync(0xFFFFFFFFu, d_hi);
            if (lane == 0) s_rowmask[tr] = ((unsigned long long)hi << 32) | lo;
            if constexpr (MASK) {
                // the mask of a tap is read at its OWN coordinates (raw cell tr + 2, tc + 2), never ring-replicated
                const unsigned char* mrow = sflag + (tr + 2) * RW + CX;
                const unsigned xlo = __ballot_sync(0xFFFFFFFFu, d_lo || !(mrow[lane] & 4));
                const unsigned xhi = __ballot_sync(0xFFFFFFFFu, (32 + lane < TW) && (d_hi || !(mrow[32 + lane] & 4)));
                if (lane == 0) s_rowexcl[tr] = ((unsigned long long)xhi << 32) | xlo;
            }
        }
    }
    __syncthreads();
    // 4b. pixels without a discontinuity in their window keep their depth; the others are COMPACTED into a list, so the
    //     expensive selection below runs in full warps (an edge crossing the tile touches a few lanes of every row-warp)
    //     A thread owns BIL_RPT CONSECUTIVE rows of one column, so the windows of its pixels overlap: each row mask is
    //     loaded and counted once (BIL_RPT + win - 1 loads instead of BIL_RPT * win) and the window counts slide.
    const int c = tile_x * BT_W + threadIdx.x;
    int n_disc_row[BIL_RPT];
    int n_excl_row[MASK ? BIL_RPT : 1];
    if constexpr (WS > 0) {
        const unsigned long long wmask = (1ull << WS) - 1ull;
        int pc[BIL_RPT + WS - 1];
#pragma unroll
        for (int q = 0; q < BIL_RPT + WS - 1; ++q)
            pc[q] = __popcll((s_rowmask[threadIdx.y * BIL_RPT + q] >> threadIdx.x) & wmask);
#pragma unroll
        for (int rr = 0; rr < BIL_RPT; ++rr) {
            int n = 0;
#pragma unroll
            for (int dr = 0; dr < WS; ++dr) n += pc[rr + dr];
            n_disc_row[rr] = n;
        }
        if constexpr (MASK) {
#pragma unroll
            for (int q = 0; q < BIL_RPT + WS - 1; ++q)
                pc[q] = __popcll((s_rowexcl[threadIdx.y * BIL_RPT + q] >> threadIdx.x) & wmask);
#pragma unroll
            for (int rr = 0; rr < BIL_RPT; ++rr) {
                int n = 0;
#pragma unroll
                for (int dr = 0; dr < WS; ++dr) n += pc[rr + dr];
                n_excl_row[rr] = n;
            }
        }
    } else {
        const unsigned long long wmask = (1ull << win) - 1ull;
#pragma unroll
        for (int rr = 0; rr < BIL_RPT; ++rr) {
            int n = 0;
            for (int dr = 0; dr < win; ++dr) n += __popcll((s_rowmask[threadIdx.y * BIL_RPT + rr + dr] >> threadIdx.x) & wmask);
            n_disc_row[rr] = n;
        }
    }
#pragma unroll
    for (int rr = 0; rr < BIL_RPT; ++rr) {
        const int ty = threadIdx.y * BIL_RPT + rr;
        const int r = tile_y * BT_H + ty;
        const bool inside = r < H && c < W;
        bool need = inside && n_disc_row[rr] > 0 && n_disc_row[rr] < win * win;
        if constexpr (MASK)  // masked pixels are skipped (:169-170); the median needs at least one tap that is neither (:188-190)
            need = inside && n_disc_row[rr] > 0 && n_excl_row[rr] < win * win && (sflag[(ty + m + 2) * RW + threadIdx.x + m + CX] & 4) != 0;
        if (inside && !need) dout[(size_t)r * W + c] = wdep[(ty + m) * wstride + threadIdx.x + m];
        const unsigned lane = tid & 31u;
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, need);
        int slot = 0;
        if (lane == 0 && mask) slot = atomicAdd(&s_count, __popc(mask));
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
        if (need) s_list[slot + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)(ty * BT_W + threadIdx.x);
    }
    __syncthreads();
    // 4c. rank-k(n) smallest among the n window pixels with disc == 0
    const int count = s_count;
    for (int idx = tid; idx < count; idx += nthr) {
        const int pix = s_list[idx];
        const int ty = pix / BT_W, tx = pix - ty * BT_W;
        const int base = ty * wstride + tx;  // top-left tap of this pixel's window
        DT result;
        if constexpr (WS > 0) {
            DT v[WS * WS];
            int n = 0;
#pragma unroll
            for (int dr = 0; dr < WS; ++dr) {
                const unsigned bits = (unsigned)((MASK ? s_rowexcl[ty + dr] : s_rowmask[ty + dr]) >> tx);
#pragma unroll
                for (int dc = 0; dc < WS; ++dc) {
                    const bool keep = !((bits >> dc) & 1u);
                    n += keep;
                    v[dr * WS + dc] = keep ? wdep[base + dr * wstride + dc] : pos_inf<DT>();
                }
            }
            network_sort<DT, WS * WS>(v);
            const int k = s_k[n];  // k <= (WS*WS - 1) / 2
            result = v[0];
#pragma unroll
            for (int q = 1; q <= (WS * WS - 1) / 2; ++q) result = (q == k) ? v[q] : result;
        } else {
            // value with  #{f < v} <= k < #{f <= v}  among the n non-discontinuity taps
            int n = 0;
            for (int dr = 0; dr < win; ++dr)
                for (int dc = 0; dc < win; ++dc) n += !wdisc[base + dr * wstride + dc];
            const int k = s_k[n];
            result = wdep[base + m * wstride + m];
            for (int er = 0; er < win; ++er) {
                for (int ec = 0; ec < win; ++ec) {
                    const int e = base + er * wstride + ec;
                    if (wdisc[e]) continue;
                    const DT v = wdep[e];
                    int lt = 0, le = 0;
                    for (int dr = 0; dr < win; ++dr)
                        for (int dc = 0; dc < win; ++dc) {
                            const int f = base + dr * wstride + dc;
                            if (!wdisc[f]) {
                                const DT u = wdep[f];
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
        dout[(size_t)(tile_y * BT_H + ty) * W + (tile_x * BT_W + tx)] = result;
    }
}

template <typename DT, int WS>
__global__ void __launch_bounds__(BT_W* BT_TY, sizeof(DT) == 4 ? OFD_BIL_MINB : 1) bilateral_iter_kernel(const DT* __restrict__ din, const DT* __restrict__ dorig,
                                                                   int H, int W, int win_rt, DT thr, DT* __restrict__ dout,
                                                                   const __grid_constant__ KRank kr) {
    bilateral_tile<DT, WS>(din, dorig, H, W, win_rt, thr, dout, blockIdx.x, blockIdx.y, kr);
}

template <int WS>
__global__ void __launch_bounds__(BT_W* BT_TY, OFD_BIL_MINB) bilateral_iter_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                                    const __grid_constant__ CUtensorMap tm_orig, int H, int W,
                                                                                    float thr, float* __restrict__ dout,
                                                                                    const __grid_constant__ KRank kr) {
    bilateral_tile<float, WS, false, true>(nullptr, nullptr, H, W, WS, thr, dout, blockIdx.x, blockIdx.y, kr, nullptr, &tm_in, &tm_orig);
}

typedef CUresult (*pfn_cuTensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static pfn_cuTensorMapEncodeTiled tensor_map_encoder() {
    static pfn_cuTensorMapEncodeTiled fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (pfn_cuTensorMapEncodeTiled)p;
    }();
    return fn;
}

// returns false when the experiment does not apply (the caller then takes the default path)
template <int WS>
static bool launch_bilateral_tma(const float* din, const float* dorig, int H, int W, float thr, float* dout, cudaStream_t st) {
    pfn_cuTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc || W % 4 != 0 || (((uintptr_t)din | (uintptr_t)dorig) & 15)) return false;
    constexpr int m = WS / 2;
    constexpr int LP = (4 - ((m + 2) & 3)) & 3;
    constexpr int RW = (BT_W + 2 * m + 4 + LP + 3) & ~3, RH = BT_H + 2 * m + 4, TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    constexpr int CELLS_P = (RW * RH + 31) & ~31;
    CUtensorMap maps[2];
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstride[1] = {(cuuint64_t)W * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)RW, (cuuint32_t)RH};
    const cuuint32_t estr[2] = {1, 1};
    const float* ptrs[2] = {din, dorig};
    for (int k = 0; k < 2; ++k)
        if (enc(&maps[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptrs[k], gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    const size_t smem = (size_t)CELLS_P * 2 * sizeof(float) + (size_t)TW * TH * (sizeof(float) + 1) + (size_t)RW * RH + 64;
    dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H), block(BT_W, BT_TY);
    ensure_dynamic_smem("bilateral_iter_tma_kernel", (const void*)bilateral_iter_tma_kernel<WS>, smem);
    bilateral_iter_tma_kernel<WS><<<grid, block, smem, st>>>(maps[0], maps[1], H, W, thr, dout, make_krank(WS));
    return true;
}

template <typename DT, int WS>
__global__ void __launch_bounds__(BT_W* BT_TY, sizeof(DT) == 4 ? OFD_BIL_MINB : 1) bilateral_masked_kernel(
    const DT* __restrict__ din, const DT* __restrict__ dorig, const unsigned char* __restrict__ mask, int H, int W, DT thr,
    DT* __restrict__ dout, const __grid_constant__ KRank kr) {
    bilateral_tile<DT, WS, true>(din, dorig, H, W, WS, thr, dout, blockIdx.x, blockIdx.y, kr, mask);
}

template <typename DT, int WS>
static void launch_bilateral_masked(const DT* din, const DT* dorig, const unsigned char* mask, int H, int W, DT thr, DT* dout,
                                    bool coef_f64, cudaStream_t st) {
    constexpr int m = WS / 2;
    constexpr int RW = BT_W + 2 * m + 4, RH = BT_H + 2 * m + 4, TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    const size_t smem = (size_t)RW * RH * (2 * sizeof(DT) + 1) + (size_t)TW * TH * (sizeof(DT) + 1) + 32;
    dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H), block(BT_W, BT_TY);
    ensure_dynamic_smem("bilateral_masked_kernel", (const void*)bilateral_masked_kernel<DT, WS>, smem);
    bilateral_masked_kernel<DT, WS><<<grid, block, smem, st>>>(din, dorig, mask, H, W, thr, dout, make_krank(WS, coef_f64));
}

template <typename DT>
static int dispatch_bilateral_masked(const DT* din, const DT* dorig, const unsigned char* mask, int H, int W, int window, DT thr,
                                     DT* dout, bool coef_f64, cudaStream_t st) {
    switch (window) {
        case 3: launch_bilateral_masked<DT, 3>(din, dorig, mask, H, W, thr, dout, coef_f64, st); return 0;
        case 5: launch_bilateral_masked<DT, 5>(din, dorig, mask, H, W, thr, dout, coef_f64, st); return 0;
        case 7: launch_bilateral_masked<DT, 7>(din, dorig, mask, H, W, thr, dout, coef_f64, st); return 0;
        default: return 1;
    }
}

// Ragged batch (BASELINE config 2: mixed-resolution frames): images of different H x W packed back to back in one
// buffer; the grid is the concatenation of every image's tile list and a CTA finds its image by bisection.
constexpr int BB_MAX = 64;
struct BilateralBatch {
    unsigned long long offset[BB_MAX];  // first element of image i in the packed buffers
    unsigned tile_start[BB_MAX + 1];    // first linear tile of image i
    int H[BB_MAX], W[BB_MAX], tiles_x[BB_MAX];
    int n;
};

template <typename DT, int WS>
__global__ void __launch_bounds__(BT_W* BT_TY, sizeof(DT) == 4 ? OFD_BIL_MINB : 1) bilateral_batch_kernel(const DT* __restrict__ din, const DT* __restrict__ dorig,
                                                                    const __grid_constant__ BilateralBatch bb, int win_rt,
                                                                    DT thr, DT* __restrict__ dout, const __grid_constant__ KRank kr) {
    const unsigned t = blockIdx.x;
    int lo = 0, hi = bb.n - 1;
    while (lo < hi) {  // last image whose tile_start <= t
        const int mid = (lo + hi + 1) >> 1;
        if (bb.tile_start[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const unsigned local = t - bb.tile_start[lo];
    const int tiles_x = bb.tiles_x[lo];
    const size_t off = (size_t)bb.offset[lo];
    bilateral_tile<DT, WS>(din + off, dorig + off, bb.H[lo], bb.W[lo], win_rt, thr, dout + off, (int)(local % tiles_x), (int)(local / tiles_x), kr);
}

template <typename DT, int WS>
static void launch_bilateral_batch(const DT* din, const DT* dorig, const BilateralBatch& bb, unsigned tiles, int window, DT thr,
                                   DT* dout, cudaStream_t st) {
    const int m = window / 2;
    const int RW = BT_W + 2 * m + 4, RH = BT_H + 2 * m + 4, TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    const size_t smem = (size_t)RW * RH * (2 * sizeof(DT) + 1) + (size_t)TW * TH * (sizeof(DT) + 1) + 32;
    ensure_dynamic_smem("bilateral_batch_kernel", (const void*)bilateral_batch_kernel<DT, WS>, smem);
    bilateral_batch_kernel<DT, WS><<<tiles, dim3(BT_W, BT_TY), smem, st>>>(din, dorig, bb, window, thr, dout, make_krank(window));
}

template <typename DT>
static void dispatch_bilateral_batch(const DT* din, const DT* dorig, const BilateralBatch& bb, unsigned tiles, int window, DT thr,
                                     DT* dout, cudaStream_t st) {
    switch (window) {
        case 3: launch_bilateral_batch<DT, 3>(din, dorig, bb, tiles, window, thr, dout, st); break;
        case 5: launch_bilateral_batch<DT, 5>(din, dorig, bb, tiles, window, thr, dout, st); break;
        case 7: launch_bilateral_batch<DT, 7>(din, dorig, bb, tiles, window, thr, dout, st); break;
        default: launch_bilateral_batch<DT, 0>(din, dorig, bb, tiles, window, thr, dout, st); break;
    }
}

template <typename DT, int WS>
static void launch_bilateral(const DT* din, const DT* dorig, int H, int W, int window, DT thr, DT* dout, cudaStream_t st) {
    const int m = window / 2;
    const int RW = BT_W + 2 * m + 4, RH = BT_H + 2 * m + 4, TW = BT_W + 2 * m, TH = BT_H + 2 * m;
    const size_t smem = (size_t)RW * RH * (2 * sizeof(DT) + 1) + (size_t)TW * TH * (sizeof(DT) + 1) + 32;
    dim3 grid((W + BT_W - 1) / BT_W, (H + BT_H - 1) / BT_H), block(BT_W, BT_TY);
    ensure_dynamic_smem("bilateral_iter_kernel", (const void*)bilateral_iter_kernel<DT, WS>, smem);
    bilateral_iter_kernel<DT, WS><<<grid, block, smem, st>>>(din, dorig, H, W, window, thr, dout, make_krank(window));
}

template <typename DT>
static void dispatch_bilateral(const DT* din, const DT* dorig, int H, int W, int window, DT thr, DT* dout, cudaStream_t st) {
    if constexpr (sizeof(DT) == 4) {
        const char* e = std::getenv("OFD_BIL_TMA");  // experiment, default off (profiles/r2/tune_bilateral_tma.txt)
        if (e && std::atoi(e) != 0) {
            bool done = false;
            if (window == 3) done = launch_bilateral_tma<3>(din, dorig, H, W, thr, dout, st);
            if (window == 5) done = launch_bilateral_tma<5>(din, dorig, H, W, thr, dout, st);
            if (window == 7) done = launch_bilateral_tma<7>(din, dorig, H, W, thr, dout, st);
            if (done) return;
        }
    }
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

extern "C" int ofd_bilateral_iter_batch(const void* depth_in, const void* depth_orig, int dtype, int n_images,
                                        const int* H_host, const int* W_host, const size_t* offset_host, int window,
                                        double threshold, void* depth_out, ofd_stream_t stream) {
    const char* fn = "ofd_bilateral_iter_batch";
    if (dtype != OFD_F32 && dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad dtype %d", fn, dtype);
    if (n_images < 0) return fail(OFD_E_SHAPE, "%s: negative image count", fn);
    if (window < 1 || window > MAX_WIN || (window & 1) == 0)
        return fail(OFD_E_ARG, "%s: window must be odd and in [1,%d]", fn, MAX_WIN);
    if (n_images == 0) return OFD_OK;
    if (!depth_in || !depth_orig || !depth_out || !H_host || !W_host || !offset_host) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    if (depth_in == depth_out) return fail(OFD_E_ARG, "%s: in-place filtering is not supported", fn);
    for (int i = 0; i < n_images; ++i)
        if (H_host[i] < 3 || W_host[i] < 3) return fail(OFD_E_SHAPE, "%s: image %d is %dx%d, needs H >= 3 and W >= 3", fn, i, H_host[i], W_host[i]);
    cudaStream_t st = (cudaStream_t)stream;
    for (int i0 = 0; i0 < n_images; i0 += BB_MAX) {
        BilateralBatch bb = {};
        bb.n = (n_images - i0) < BB_MAX ? (n_images - i0) : BB_MAX;
        unsigned long long tiles = 0;
        for (int k = 0; k < bb.n; ++k) {
            const int H = H_host[i0 + k], W = W_host[i0 + k];
            bb.H[k] = H, bb.W[k] = W, bb.offset[k] = offset_host[i0 + k];
            bb.tiles_x[k] = (W + BT_W - 1) / BT_W;
            bb.tile_start[k] = (unsigned)tiles;
            tiles += (unsigned long long)bb.tiles_x[k] * ((H + BT_H - 1) / BT_H);
        }
        bb.tile_start[bb.n] = (unsigned)tiles;
        if (tiles > 0x7FFFFFFFull) return fail(OFD_E_SHAPE, "%s: too many tiles in one launch", fn);
        if (dtype == OFD_F32)
            dispatch_bilateral_batch<float>((const float*)depth_in, (const float*)depth_orig, bb, (unsigned)tiles, window, (float)threshold,
                                            (float*)depth_out, st);
        else
            dispatch_bilateral_batch<double>((const double*)depth_in, (const double*)depth_orig, bb, (unsigned)tiles, window, threshold,
                                             (double*)depth_out, st);
        int rc = check_launch(fn);
        if (rc) return rc;
    }
    return OFD_OK;
}

extern "C" int ofd_bilateral_iter_masked(const void* depth_in, const void* depth_orig, const uint8_t* mask, int mask_coef_f64,
                                         int dtype, int H, int W, int window, double threshold, void* depth_out,
                                         ofd_stream_t stream) {
    const char* fn = "ofd_bilateral_iter_masked";
    if (dtype != OFD_F32 && dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad dtype %d", fn, dtype);
    if (H < 3 || W < 3) return fail(OFD_E_SHAPE, "%s: needs H >= 3 and W >= 3", fn);
    if ((H + BT_H - 1) / BT_H > 65535) return fail(OFD_E_SHAPE, "%s: H too large", fn);
    if (window != 3 && window != 5 && window != 7) return fail(OFD_E_ARG, "%s: the mask path supports windows 3, 5 and 7", fn);
    if (!depth_in || !depth_orig || !mask || !depth_out) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    if (depth_in == depth_out) return fail(OFD_E_ARG, "%s: in-place filtering is not supported", fn);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OFD_F32)
        dispatch_bilateral_masked<float>((const float*)depth_in, (const float*)depth_orig, mask, H, W, window, (float)threshold,
                                         (float*)depth_out, mask_coef_f64 != 0, st);
    else
        dispatch_bilateral_masked<double>((const double*)depth_in, (const double*)depth_orig, mask, H, W, window, threshold,
                                          (double*)depth_out, mask_coef_f64 != 0, st);
    return check_launch(fn);
}
