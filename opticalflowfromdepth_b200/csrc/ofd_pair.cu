// ofd_pair.cu — one whole virtual-stereo "flow pair" (preprocess.py:355-366 minus inpaint) in one kernel.
//
//   disp  = sBf / depth0                      Convert.depth_to_disparity   preprocess.py:239-246
//   flow  = (-disp, -0.0)                     Convert.disparity_to_flow    preprocess.py:249-254
//   obj   = img0(3) | depth0(1) | -flow(2)    preprocess.py:358
//   splat obj along flow, z-test depth0       alt_cuda/fw.py:19-59 + fw_cuda_kernel.cu:28-47
//   outputs * valid, fix_warped_depth         preprocess.py:362-365, utils.py:123-126
//
// flow.y is exactly -0.0, so (float)j + flow.y == j: no source ever leaves its row.  One CTA owns one row and
// keeps that row's z-buffer in shared memory: two native 32-bit shared-memory atomicMin passes (min ordered
// depth, then min source column among the depth-minimal sources) reproduce the serial loop's winner exactly,
// with no global atomics and no key workspace.  The input row (depth + 3 colour planes) is brought in by the
// TMA engine as 1-D bulk copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers, so the colour planes land
// while the z-test runs.  HBM traffic is the algorithmic minimum: 16 B/px in, 40 B/px out.
#include <cstdio>
#include <cstdlib>

#include "ofd_common.cuh"

namespace ofd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory carve-up for a row of W pixels (DT = depth dtype)
template <typename DT>
struct PairSmem {
    static __host__ __device__ size_t bytes(int W) {
        // raw depth row | img rows (3) | sdepth | sdisp | target | ord | idx | 2 mbarriers
        return 16 + (size_t)W * (sizeof(DT) + 3 * 4 + 4 + 4 + 4 + 4 + 4) + 64;
    }
};

template <typename DT, bool BULK>
__global__ void __launch_bounds__(256) pair_row_kernel(const float* __restrict__ img0, const DT* __restrict__ depth0,
                                                      const float* __restrict__ sBf, float* __restrict__ img1,
                                                      float* __restrict__ depth1, float* __restrict__ back_flow,
                                                      float* __restrict__ flow, float* __restrict__ valid,
                                                      float* __restrict__ collision, uint64_t* __restrict__ counters,
                                                      int H, int W) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int j = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t hw = (size_t)H * W;
    const size_t row = (size_t)j * W;

    // carve (every region starts 16-byte aligned when W % 4 == 0; the bulk path requires that)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [0] depth, [1] colour
    DT* sraw = reinterpret_cast<DT*>(smem_raw + 16);
    float* simg = reinterpret_cast<float*>(sraw + W);
    float* sdepth = simg + 3 * (size_t)W;
    float* sdisp = sdepth + W;
    uint32_t* stgt = reinterpret_cast<uint32_t*>(sdisp + W);
    uint32_t* sord = stgt + W;
    uint32_t* sidx = sord + W;

    const float* img_b = img0 + (size_t)b * 3 * hw + row;
    const DT* dep_b = depth0 + (size_t)b * hw + row;

    if (BULK) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bars[0], (unsigned)(W * sizeof(DT)));
            bulk_g2s(sraw, dep_b, (unsigned)(W * sizeof(DT)), &bars[0]);
            mbar_expect_tx(&bars[1], (unsigned)(3 * W * 4));
            for (int c = 0; c < 3; ++c) bulk_g2s(simg + (size_t)c * W, img_b + c * hw, (unsigned)(W * 4), &bars[1]);
        }
    } else {
        for (int i = tid; i < W; i += nt) {
            sraw[i] = dep_b[i];
            for (int c = 0; c < 3; ++c) simg[(size_t)c * W + i] = img_b[c * hw + i];
        }
    }
    for (int i = tid; i < W; i += nt) {
        sord[i] = 0xFFFFFFFFu;
        sidx[i] = 0xFFFFFFFFu;
    }
    if (BULK)
        mbar_wait(&bars[0], 0);
    else
        __syncthreads();

    // ---- phase A: disparity, flow, target column ------------------------------------------------------
    const DT s = (DT)sBf[b];  // float32 scalar promoted to the depth dtype (0-dim tensor rule)
    float* flow_b = flow ? flow + (size_t)b * 2 * hw + row : nullptr;
    for (int i = tid; i < W; i += nt) {
        const DT d = sraw[i];
        const DT disp = s / d;
        const DT fx = disp * (DT)-1.0;
        uint32_t tx = T_DROPPED;
        DT px = (DT)(float)i + fx;
        if (px == px) {
            px = px < (DT)0 ? (DT)0 : px;
            px = px > (DT)(W - 1) ? (DT)(W - 1) : px;
            tx = (uint32_t)(int)px;
        }
        stgt[i] = tx;
        sdepth[i] = (float)d;
        sdisp[i] = (float)(fx * (DT)-1.0);
        if (flow_b) {
            flow_b[i] = (float)fx;
            flow_b[hw + i] = -0.0f;
        }
    }
    __syncthreads();
    // ---- phase B: min ordered depth per target ----------------------------------------------------------
    unsigned dropped = 0;
    for (int i = tid; i < W; i += nt) {
        const uint32_t tx = stgt[i];
        if (tx != T_DROPPED)
            atomicMin(&sord[tx], depth_hi(sdepth[i]));
        else
            dropped++;
    }
    __syncthreads();
    // ---- phase C: lowest source column among the depth-minimal sources -----------------------------------
    for (int i = tid; i < W; i += nt) {
        const uint32_t tx = stgt[i];
        if (tx != T_DROPPED && sord[tx] == depth_hi(sdepth[i])) atomicMin(&sidx[tx], (uint32_t)i);
    }
    __syncthreads();
    if (BULK) mbar_wait(&bars[1], 0);

    // ---- phase D: gather + epilogue ------------------------------------------------------------------------
    float* img1_b = img1 + (size_t)b * 3 * hw + row;
    float* dep1_b = depth1 + (size_t)b * hw + row;
    float* bf_b = back_flow + (size_t)b * 2 * hw + row;
    float* val_b = valid + (size_t)b * hw + row;
    float* col_b = collision ? collision + (size_t)b * hw + row : nullptr;
    unsigned n_hit = 0, n_col = 0, n_px = 0;
    for (int t = tid; t < W; t += nt) {
        const uint32_t hi = sord[t];
        const bool hit = hi != 0xFFFFFFFFu;
        const bool win = hi < HI_NOWIN;
        const uint32_t src = win ? sidx[t] : 0u;
        const float v = hit ? 1.0f : 0.0f;
        float r = 0.f, g = 0.f, bl = 0.f, dd = 0.f, bx = 0.f;
        if (win) {
            r = simg[src];
            g = simg[(size_t)W + src];
            bl = simg[2 * (size_t)W + src];
            dd = sdepth[src];
            bx = sdisp[src];
        }
        img1_b[t] = r * v;
        img1_b[hw + t] = g * v;
        img1_b[2 * hw + t] = bl * v;
        dep1_b[t] = fix_depth(dd * v);
        bf_b[t] = bx * v;
        bf_b[hw + t] = 0.0f;  // (-0.0 * -1.0) * valid
        val_b[t] = v;
        if (col_b) col_b[t] = (hit && !win) ? 1.0f : 0.0f;
        n_px++;
        n_hit += hit;
        n_col += (hit && !win);
    }
    if (counters) {
        unsigned n_tie = 0;
        for (int i = tid; i < W; i += nt) {
            const uint32_t tx = stgt[i], hi = depth_hi(sdepth[i]);
            if (tx != T_DROPPED && hi < HI_NOWIN && sord[tx] == hi && sidx[tx] != (uint32_t)i) n_tie++;
        }
        warp_count(counters, OFD_CNT_HIT, n_hit);
        warp_count(counters, OFD_CNT_HOLE, n_px - n_hit);
        warp_count(counters, OFD_CNT_COLLISION, n_col);
        warp_count(counters, OFD_CNT_DROPPED, dropped);
        warp_count(counters, OFD_CNT_TIE_SRC, n_tie);
    }
}

// ---- v2: persistent, TMA-in / TMA-out row pipeline ---------------------------------------------------------------
// ncu on the one-row-per-CTA kernel above showed it ISSUE-bound (sm__throughput 84 %, ~330 instructions per pixel,
// mostly 64-bit global address arithmetic and scalar stores), not HBM-bound.  Here every global access is a TMA
// 1-D bulk copy (UBLKCP): input rows are prefetched up to three work units ahead into an mbarrier ring (2-4 stages, chosen per
// row width), results are staged in a 1-3 deep shared-memory ring and written back by bulk stores that drain while the next
// units are computed; the SIMT threads only touch shared memory (32-bit addressing, no per-pixel global pointer math).  Each
// thread owns NITER fixed pixels of the (virtual) row and keeps their target column and ordered depth in registers across
// the three z-buffer phases.
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef OFD_PAIR_NITER
#define OFD_PAIR_NITER 2
#endif
constexpr int kMaxInStages = 6;  // input ring depth is a launch parameter: rows are prefetched in_stages-1 ahead

// shared-memory plan of the persistent kernel, in floats after a 64-byte header (mbarriers)
template <typename DT>
struct PairPlan {
    static constexpr int kIn = (int)(sizeof(DT) / 4) + 3;  // raw depth row (DT) + 3 colour rows, per input stage
    static constexpr int kOut = 8;                          // img1 x3, depth1, back_flow.x, flow.x, valid, collision
    static constexpr int kMisc = 4 + (sizeof(DT) == 8 ? 1 : 0);  // zero row, -0 row, ord, idx (+ float depth row)
    static __host__ __device__ size_t bytes(int W, int in_stages, int out_stages) {
        return 128 + (size_t)W * 4 * (in_stages * kIn + out_stages * kOut + kMisc);
    }
};

// Ragged batches (BASELINE config 2: mixed-resolution frames): image i is H x W, its planes start at pixel offset `off` of the
// packed buffers (a C-channel tensor stores it densely at element C * off), it is cut into units of G rows and its first unit has
// the batch-wide index unit0.  The table travels in the kernel's parameter space: every thread of a CTA reads the same entry.
constexpr int kPairRaggedMax = 96;  // images per launch
struct PairRaggedImg {
    int H, W, G, unit0;
    unsigned long long off;
};
struct PairRaggedTable {
    int n, total_units;
    PairRaggedImg img[kPairRaggedMax + 1];  // img[n].unit0 == total_units (sentinel)
};
struct PairNoTable {};
// PAIR_RAGGED: every frame's H*W and offset are multiples of 4 pixels (all planes of a unit share one 16-byte phase);
// PAIR_RAGGED_ANY: no such rule, every plane has its own phase (more registers and per-plane store bookkeeping: only when needed)
enum { PAIR_ROW = 0, PAIR_GROUPED = 1, PAIR_RAGGED = 2, PAIR_RAGGED_ANY = 3 };
template <int MODE>
struct PairTab {
    typedef PairNoTable type;
};
template <>
struct PairTab<PAIR_RAGGED> {
    typedef PairRaggedTable type;
};
template <>
struct PairTab<PAIR_RAGGED_ANY> {
    typedef PairRaggedTable type;
};

template <typename DT, int NITER, int MODE>
__global__ void __launch_bounds__(1024) pair_rows_persistent(const float* __restrict__ img0, const DT* __restrict__ depth0,
                                                           const float* __restrict__ sBf, float* __restrict__ img1,
                                                           float* __restrict__ depth1, float* __restrict__ back_flow,
                                                           float* __restrict__ flow, float* __restrict__ valid,
                                                           float* __restrict__ collision, uint64_t* __restrict__ counters,
                                                           int B, int H, int W, int G, int in_stages_rt, int out_stages_rt,
                                                           const __grid_constant__ typename PairTab<MODE>::type tab) {
    constexpr bool GROUPED = MODE != PAIR_ROW;
    constexpr bool RAGGED = MODE >= PAIR_RAGGED;  // H is unused, W = the widest unit of the batch (shared-memory row stride), G = 1
    constexpr bool ANY = MODE == PAIR_RAGGED_ANY;
    // A work unit is G consecutive rows of one frame, handled as ONE virtual row of VW = G * W pixels: the rows are contiguous in
    // every plane, so each plane still moves with one bulk copy per unit; sources stay in their row, so the row-local z-buffer
    // only needs the virtual target index r * W + tx.  G > 1 (a) makes narrow rows long enough to amortise the per-unit barrier /
    // TMA-issue chain and (b) restores the 16-byte size / address alignment TMA needs when W is not a multiple of 4 (G even for
    // W % 4 == 2, G % 4 == 0 for odd W; the launcher only takes this kernel when H is a multiple of that alignment).
#ifdef OFD_PAIR_IN_STAGES
    constexpr int kInStages = OFD_PAIR_IN_STAGES, kOutStages = OFD_PAIR_OUT_STAGES;  // tuning builds: compile-time rings
#else
    const int kInStages = in_stages_rt, kOutStages = out_stages_rt;
#endif
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t hw = (size_t)H * W;
    // GROUPED = false is the G == 1 specialisation: the index arithmetic below folds back to the one-row form
    const int VW = (GROUPED && !RAGGED) ? G * W : W;
    const int upf = RAGGED ? 1 : (GROUPED ? (H + G - 1) / G : H);  // units per frame
    int total_rows = B * upf;                                     // work units
    if constexpr (RAGGED) total_rows = tab.total_units;
    // where a unit lives: frame, first row, pixels, row width, pixel offset and plane size of its frame
    struct Unit {
        int b, j, NV, W;
        size_t off, hw;
    };
    auto locate = [&](int row, int& cur) -> Unit {
        Unit u;
        if constexpr (RAGGED) {
            while (row >= tab.img[cur + 1].unit0) ++cur;  // a CTA's units only move forward
            const PairRaggedImg& t = tab.img[cur];
            u.b = cur, u.W = t.W, u.j = (row - t.unit0) * t.G;
            u.NV = ((t.H - u.j) < t.G ? (t.H - u.j) : t.G) * t.W;
            u.off = (size_t)t.off, u.hw = (size_t)t.H * t.W;
        } else {
            u.b = row / upf, u.j = GROUPED ? (row - u.b * upf) * G : row - u.b * upf;
            u.W = W, u.NV = GROUPED ? ((H - u.j) < G ? (H - u.j) : G) * W : W;
            u.hw = hw, u.off = (size_t)u.b * hw;
        }
        return u;
    };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [stage][0 depth, 1 colour]
    float* base = reinterpret_cast<float*>(smem_raw + 128);
    typedef PairPlan<DT> Plan;
    auto in_stage = [&](int s) { return base + (size_t)s * Plan::kIn * VW; };
    auto out_stage = [&](int s) { return base + (size_t)(kInStages * Plan::kIn + s * Plan::kOut) * VW; };
    float* misc = base + (size_t)(kInStages * Plan::kIn + kOutStages * Plan::kOut) * VW;
    float* zero_row = misc;
    float* negzero_row = misc + VW;
    uint32_t* sord = reinterpret_cast<uint32_t*>(misc + 2 * (size_t)VW);
    uint32_t* sidx = sord + VW;
    float* sdepth32 = misc + 4 * (size_t)VW;  // only when DT is double

    if (tid == 0) {
        for (int k = 0; k < 2 * kInStages; ++k) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < VW; i += nt) {
        zero_row[i] = 0.0f;
        negzero_row[i] = -0.0f;
    }
    fence_async_smem();
    __syncthreads();
    // the thread's pixels are the same virtual indices in every unit: row-in-unit and column are computed once
    // (ragged batches: recomputed whenever the CTA moves on to another frame)
    int vr[NITER], vi[NITER];
#pragma unroll
    for (int k = 0; k < NITER; ++k) {
        const int v = tid + k * nt;
        vr[k] = (GROUPED && !RAGGED) ? v / W : 0;
        vi[k] = (GROUPED && !RAGGED) ? v - vr[k] * W : v;
    }

    constexpr int kDPer = 16 / (int)sizeof(DT);  // depth elements per 16 bytes
    int cur_ld = 0, cur = 0, cur_w = -1;  // table cursors of the loader / of the CTA, frame the vr / vi are valid for
    auto issue_loads = [&](int row, int s) {  // one thread (the loader)
        const Unit u = locate(row, cur_ld);
        const int n = u.NV;  // pixels of this unit
        DT* sraw = reinterpret_cast<DT*>(in_stage(s));
        float* simg = in_stage(s) + (sizeof(DT) / 4) * (size_t)VW;
        // ragged units may start anywhere: the copy covers the 16-byte-aligned superset of the unit (the extra elements belong to
        // the neighbouring rows of the same plane: plane sizes and offsets are multiples of 4 pixels), pixel i sits at [i + shift]
        if constexpr (ANY) {
            // every plane has its own 16-byte phase (a frame's planes are H*W elements apart, not necessarily a multiple of 4)
            const size_t q0 = (size_t)u.j * u.W, e1 = u.off + q0, e3 = 3 * u.off + q0;
            const int ad = (int)(e1 & (kDPer - 1));
            const unsigned dbytes = (unsigned)(((ad + n + kDPer - 1) & ~(kDPer - 1)) * sizeof(DT));
            mbar_expect_tx(&bars[2 * s], dbytes);
            bulk_g2s(sraw, depth0 + e1 - ad, dbytes, &bars[2 * s]);
            unsigned cb[3], total = 0;
            for (int c = 0; c < 3; ++c) cb[c] = (unsigned)((((int)((e3 + c * u.hw) & 3) + n + 3) & ~3) * 4), total += cb[c];
            mbar_expect_tx(&bars[2 * s + 1], total);
            for (int c = 0; c < 3; ++c) {
                const size_t e = e3 + c * u.hw;
                bulk_g2s(simg + (size_t)c * VW, img0 + (e - (e & 3)), cb[c], &bars[2 * s + 1]);
            }
        } else if constexpr (RAGGED) {
            const size_t e0 = u.off + (size_t)u.j * u.W;
            const int a = (int)(e0 & 3), ad = (int)(e0 & (kDPer - 1));
            const unsigned dbytes = (unsigned)(((ad + n + kDPer - 1) & ~(kDPer - 1)) * sizeof(DT));
            const unsigned cbytes = (unsigned)(((a + n + 3) & ~3) * 4);
            mbar_expect_tx(&bars[2 * s], dbytes);
            bulk_g2s(sraw, depth0 + e0 - ad, dbytes, &bars[2 * s]);
            mbar_expect_tx(&bars[2 * s + 1], 3 * cbytes);
            const float* g = img0 + 3 * u.off + (size_t)u.j * u.W - a;
            for (int c = 0; c < 3; ++c) bulk_g2s(simg + (size_t)c * VW, g + c * u.hw, cbytes, &bars[2 * s + 1]);
        } else {
            mbar_expect_tx(&bars[2 * s], (unsigned)(n * sizeof(DT)));
            bulk_g2s(sraw, depth0 + u.off + (size_t)u.j * u.W, (unsigned)(n * sizeof(DT)), &bars[2 * s]);
            mbar_expect_tx(&bars[2 * s + 1], (unsigned)(3 * n * 4));
            const float* g = img0 + 3 * u.off + (size_t)u.j * u.W;
            for (int c = 0; c < 3; ++c) bulk_g2s(simg + (size_t)c * VW, g + c * u.hw, (unsigned)(n * 4), &bars[2 * s + 1]);
        }
    };

    int row = blockIdx.x;
    const int loader = nt > 32 ? 32 : 0;  // loads are issued by warp 1 so they do not queue behind warp 0's store bookkeeping
    if (tid == loader)
        for (int k = 0; k < kInStages - 1; ++k)
            if (row + k * (int)gridDim.x < total_rows) issue_loads(row + k * (int)gridDim.x, k);
    unsigned n_hit = 0, n_col = 0, n_px = 0, n_drop = 0, n_tie = 0;

    // ring cursors advance incrementally (no runtime division in the per-row critical path)
    int s = 0, so = 0, s_fill = kInStages - 1;
    unsigned ph = 0;
    for (; row < total_rows; row += gridDim.x, s = (s + 1 == kInStages ? 0 : s + 1), so = (so + 1 == kOutStages ? 0 : so + 1),
                             s_fill = (s_fill + 1 == kInStages ? 0 : s_fill + 1), ph ^= (s == 0 ? 1u : 0u)) {
        const Unit u = locate(row, cur);
        const int b = u.b, j = u.j;
        const int NV = u.NV;  // pixels of this unit (the last unit of a frame may be shorter)
        const int Wu = u.W;
        int a = 0, ad = 0;  // ragged units: shift of the float rows / of the raw depth row in shared memory (see issue_loads)
        if constexpr (RAGGED) {
            const size_t e0 = u.off + (size_t)j * Wu;  // first pixel of the unit in its planes
            a = (int)(e0 & 3), ad = (int)(e0 & (kDPer - 1));
        }
        // per-plane shifts (PAIR_RAGGED_ANY; otherwise all equal to a): a3x colour plane x, a2x / a2y the x / y planes of the flows
        int a30 = a, a31 = a, a32 = a, a2x = a, a2y = a;
        if constexpr (ANY) {
            const size_t q0 = (size_t)j * Wu;
            a30 = (int)((3 * u.off + q0) & 3), a31 = (int)((3 * u.off + u.hw + q0) & 3), a32 = (int)((3 * u.off + 2 * u.hw + q0) & 3);
            a2x = (int)((2 * u.off + q0) & 3), a2y = (int)((2 * u.off + u.hw + q0) & 3);
        }
        const int as = sizeof(DT) == 4 ? a : 0;  // shift of the float depth row (sdepth32 is unshifted)
        if constexpr (RAGGED) {
            if (b != cur_w) {
                cur_w = b;
#pragma unroll
                for (int k = 0; k < NITER; ++k) {
                    const int v = tid + k * nt;
                    vr[k] = v / Wu;
                    vi[k] = v - vr[k] * Wu;
                }
            }
        }
        const DT* sraw = reinterpret_cast<const DT*>(in_stage(s));
        const float* simg = in_stage(s) + (sizeof(DT) / 4) * (size_t)VW;
        const float* sdep = sizeof(DT) == 4 ? reinterpret_cast<const float*>(sraw) : sdepth32;
        float* o_img = out_stage(so);
        float* o_dep = o_img + 3 * (size_t)VW;
        float* o_bfx = o_dep + VW;
        float* o_flx = o_bfx + VW;
        float* o_val = o_flx + VW;
        float* o_col = o_val + VW;

        if (tid == loader) {
            // the stage being refilled was released by the barrier that ended row n-1
            const int next = row + (kInStages - 1) * (int)gridDim.x;
            if (next < total_rows) issue_loads(next, s_fill);
        }
        if (tid == 0) {
            // the stores that last used out_stage(so) have finished reading it
            if (kOutStages >= 3)
                bulk_wait_read<2>();
            else if (kOutStages == 2)
                bulk_wait_read<1>();
            else
                bulk_wait_read<0>();
        }
#pragma unroll
        for (int k = 0; k < NITER; ++k) {
            const int i = tid + k * nt;
            if (i < NV) sord[i] = 0xFFFFFFFFu, sidx[i] = 0xFFFFFFFFu;
        }
        mbar_wait(&bars[2 * s], ph);
        __syncthreads();

        // ---- phase A: disparity, flow, target column, min ordered depth ------------------------------------------
        const DT sc = (DT)sBf[b];
        uint32_t tx[NITER], hi[NITER];
#pragma unroll
        for (int k = 0; k < NITER; ++k) {
            const int i = tid + k * nt;  // virtual index: row vr[k] of the unit, column vi[k]
            tx[k] = T_DROPPED;
            hi[k] = 0;
            if (i < NV) {
                const DT d = sraw[i + ad];
                const DT fx = (sc / d) * (DT)-1.0;
                DT px = (DT)(float)vi[k] + fx;
                if (px == px) {
                    px = px < (DT)0 ? (DT)0 : px;
                    px = px > (DT)(Wu - 1) ? (DT)(Wu - 1) : px;
                    tx[k] = GROUPED ? (uint32_t)(vr[k] * Wu + (int)px) : (uint32_t)(int)px;
                }
                o_flx[i + a2x] = (float)fx;
                if (sizeof(DT) == 8) sdepth32[i] = (float)d;
                hi[k] = depth_hi((float)d);
                if (tx[k] != T_DROPPED)
                    atomicMin(&sord[tx[k]], hi[k]);
                else
                    n_drop++;
            }
        }
        __syncthreads();
        // ---- phase C: lowest source column among the depth-minimal sources ---------------------------------------
#pragma unroll
        for (int k = 0; k < NITER; ++k)
            if (tx[k] != T_DROPPED && sord[tx[k]] == hi[k]) atomicMin(&sidx[tx[k]], (uint32_t)(tid + k * nt));
        mbar_wait(&bars[2 * s + 1], ph);
        __syncthreads();
        // ---- phase D: gather + epilogue into the staging rows ------------------------------------------------------
#pragma unroll
        for (int k = 0; k < NITER; ++k) {
            const int t = tid + k * nt;
            if (t < NV) {
                const uint32_t h32 = sord[t];
                const bool hit = h32 != 0xFFFFFFFFu;
                const bool win = h32 < HI_NOWIN;
                const uint32_t src = win ? sidx[t] : 0u;
                const float v = hit ? 1.0f : 0.0f;
                float r = 0.f, g = 0.f, bl = 0.f, dd = 0.f, bx = 0.f;
                if (win) {
                    r = simg[src + a30];
                    g = simg[VW + src + a31];
                    bl = simg[2 * VW + src + a32];
                    dd = sdep[src + as];
                    bx = o_flx[src + a2x] * -1.0f;
                }
                o_img[t + a30] = r * v;
                o_img[VW + t + a31] = g * v;
                o_img[2 * VW + t + a32] = bl * v;
                o_dep[t + a] = fix_depth(dd * v);
                o_bfx[t + a2x] = bx * v;
                o_val[t + a] = v;
                o_col[t + a] = (hit && !win) ? 1.0f : 0.0f;
                n_px++;
                n_hit += hit;
                n_col += (hit && !win);
            }
        }
        if (counters) {  // tie census: sources that tie the winning depth of their target but lost on column order
#pragma unroll
            for (int k = 0; k < NITER; ++k)
                if (tx[k] != T_DROPPED && hi[k] < HI_NOWIN && sord[tx[k]] == hi[k] && sidx[tx[k]] != (uint32_t)(tid + k * nt)) n_tie++;
        }
        fence_async_smem();  // generic-proxy writes of the staging rows -> visible to the TMA engine
        __syncthreads();
        if constexpr (!RAGGED) {
            if (tid == 0) {
                const size_t r1 = u.off + (size_t)j * Wu;
                const unsigned rb = (unsigned)(NV * 4);
                float* gi = img1 + 3 * u.off + (size_t)j * Wu;
                for (int c = 0; c < 3; ++c) bulk_s2g(gi + c * u.hw, o_img + (size_t)c * VW, rb);
                bulk_s2g(depth1 + r1, o_dep, rb);
                float* gb = back_flow + 2 * u.off + (size_t)j * Wu;
                bulk_s2g(gb, o_bfx, rb);
                bulk_s2g(gb + u.hw, zero_row, rb);
                if (flow) {
                    float* gf = flow + 2 * u.off + (size_t)j * Wu;
                    bulk_s2g(gf, o_flx, rb);
                    bulk_s2g(gf + u.hw, negzero_row, rb);
                }
                bulk_s2g(valid + r1, o_val, rb);
                if (collision) bulk_s2g(collision + r1, o_col, rb);
                bulk_commit();
            }
        }
        if constexpr (ANY) {
            // as below, but every plane has its own shift, hence its own head / interior / tail.
            // plane p: 0-2 img1, 3 depth1, 4 back_flow.x, 5 back_flow.y, 6 flow.x, 7 flow.y, 8 valid, 9 collision
            auto plane = [&](int p, float*& dst, const float*& src, int& sh) {
                const size_t q0 = (size_t)j * Wu;
                if (p < 3) dst = img1 + 3 * u.off + p * u.hw + q0, src = o_img + (size_t)p * VW, sh = p == 0 ? a30 : (p == 1 ? a31 : a32);
                else if (p == 3) dst = depth1 + u.off + q0, src = o_dep, sh = a;
                else if (p == 4) dst = back_flow + 2 * u.off + q0, src = o_bfx, sh = a2x;
                else if (p == 5) dst = back_flow + 2 * u.off + u.hw + q0, src = nullptr, sh = a2y;
                else if (p == 6) dst = flow ? flow + 2 * u.off + q0 : nullptr, src = o_flx, sh = a2x;
                else if (p == 7) dst = flow ? flow + 2 * u.off + u.hw + q0 : nullptr, src = nullptr, sh = a2y;
                else if (p == 8) dst = valid + u.off + q0, src = o_val, sh = a;
                else dst = collision ? collision + u.off + q0 : nullptr, src = o_col, sh = a;
            };
            if (tid == 0) {
#pragma unroll
                for (int p = 0; p < 10; ++p) {
                    float* dst;
                    const float* src;
                    int sh;
                    plane(p, dst, src, sh);
                    const int head = ((4 - sh) & 3) < NV ? ((4 - sh) & 3) : NV;
                    const int mid = (NV - head) & ~3;
                    if (dst && mid > 0) {
                        const float* from = src ? src + sh + head : (p == 5 ? zero_row : negzero_row);  // sh + head is 0 or 4
                        bulk_s2g(dst + head, from, (unsigned)(mid * 4));
                    }
                }
                bulk_commit();  // a unit without an aligned interior still commits its (empty) group: the ring counts groups
            }
            if (tid >= 32 && tid < 92) {  // (warp 0 is busy issuing the bulk stores)
                const int pl = (tid - 32) / 6, sl = (tid - 32) - pl * 6;  // plane 0-9, slot: 0-2 head, 3-5 tail
                float* dst;
                const float* src;
                int sh;
                plane(pl, dst, src, sh);
                const int head = ((4 - sh) & 3) < NV ? ((4 - sh) & 3) : NV;
                const int mid = (NV - head) & ~3;
                const int tail = NV - head - mid;
                const int t = sl < 3 ? (sl < head ? sl : -1) : (sl - 3 < tail ? head + mid + sl - 3 : -1);
                if (t >= 0 && dst) dst[t] = src ? src[t + sh] : (pl == 5 ? 0.0f : -0.0f);
            }
        } else if constexpr (RAGGED) {
            // bulk stores cover the 16-byte-aligned interior of the unit: `head` pixels before it and `tail` pixels after it (at
            // most 3 each) are written by ordinary stores of the first threads
            const int head = ((4 - a) & 3) < NV ? ((4 - a) & 3) : NV;
            const int mid = (NV - head) & ~3;
            const int tail = NV - head - mid;
            if (tid == 0) {
                if (mid > 0) {
                    const size_t r1 = u.off + (size_t)j * Wu + head;
                    const int sh = a + head;  // aligned: a + head is 0 or 4
                    const unsigned rb = (unsigned)(mid * 4);
                    float* gi = img1 + 3 * u.off + (size_t)j * Wu + head;
                    for (int c = 0; c < 3; ++c) bulk_s2g(gi + c * u.hw, o_img + (size_t)c * VW + sh, rb);
                    bulk_s2g(depth1 + r1, o_dep + sh, rb);
                    float* gb = back_flow + 2 * u.off + (size_t)j * Wu + head;
                    bulk_s2g(gb, o_bfx + sh, rb);
                    bulk_s2g(gb + u.hw, zero_row, rb);
                    if (flow) {
                        float* gf = flow + 2 * u.off + (size_t)j * Wu + head;
                        bulk_s2g(gf, o_flx + sh, rb);
                        bulk_s2g(gf + u.hw, negzero_row, rb);
                    }
                    bulk_s2g(valid + r1, o_val + sh, rb);
                    if (collision) bulk_s2g(collision + r1, o_col + sh, rb);
                }
                bulk_commit();  // a unit without an aligned interior still commits its (empty) group: the ring counts groups
            }
            if ((head | tail) != 0 && tid < 60) {
                const int pl = tid / 6, sl = tid - pl * 6;  // plane 0-9, slot: 0-2 head, 3-5 tail
                const int t = sl < 3 ? (sl < head ? sl : -1) : (sl - 3 < tail ? head + mid + sl - 3 : -1);
                if (t >= 0) {
                    const size_t px = (size_t)j * Wu + t;
                    float* dst;
                    float val;
                    if (pl < 3) dst = img1 + 3 * u.off + pl * u.hw + px, val = o_img[(size_t)pl * VW + t + a];
                    else if (pl == 3) dst = depth1 + u.off + px, val = o_dep[t + a];
                    else if (pl == 4) dst = back_flow + 2 * u.off + px, val = o_bfx[t + a];
                    else if (pl == 5) dst = back_flow + 2 * u.off + u.hw + px, val = 0.0f;
                    else if (pl == 6) dst = flow ? flow + 2 * u.off + px : nullptr, val = o_flx[t + a];
                    else if (pl == 7) dst = flow ? flow + 2 * u.off + u.hw + px : nullptr, val = -0.0f;
                    else if (pl == 8) dst = valid + u.off + px, val = o_val[t + a];
                    else dst = collision ? collision + u.off + px : nullptr, val = o_col[t + a];
                    if (dst) *dst = val;
                }
            }
        }
    }
    if (tid == 0) bulk_wait_all();
    if (counters) {
        warp_count(counters, OFD_CNT_HIT, n_hit);
        warp_count(counters, OFD_CNT_HOLE, n_px - n_hit);
        warp_count(counters, OFD_CNT_COLLISION, n_col);
        warp_count(counters, OFD_CNT_DROPPED, n_drop);
        warp_count(counters, OFD_CNT_TIE_SRC, n_tie);
    }
}

// Ring depths for a shared-memory row stride of VW pixels, tuned on B200 (profiles/r1/tune_pair.txt): the deepest plan that still lets
// two CTAs share an SM, else the deepest that fits one.  false: the row is too wide for shared memory.
template <typename DT>
static bool choose_pair_plan(int VW, int* in_stages, int* out_stages, size_t* smem) {
#ifdef OFD_PAIR_IN_STAGES
    static const int plans[][2] = {{OFD_PAIR_IN_STAGES, OFD_PAIR_OUT_STAGES}};
#else
    static const int plans[][2] = {{4, 3}, {3, 3}, {3, 2}, {2, 2}, {2, 1}};
#endif
    for (int pass = 0; pass < 2; ++pass)
        for (const auto& pl : plans) {
            const size_t need = PairPlan<DT>::bytes(VW, pl[0], pl[1]);
            if (need <= (pass == 0 ? (size_t)113 * 1024 : (size_t)227 * 1024)) {
                *in_stages = pl[0], *out_stages = pl[1], *smem = need;
                return true;
            }
        }
    return false;
}

template <typename DT>
static int launch_pair_persistent(const char* fn, const float* img0, const DT* depth0, const float* sBf, int B, int H, int W,
                                  float* img1, float* depth1, float* back_flow, float* flow, float* valid, float* collision,
                                  uint64_t* counters, cudaStream_t st, bool* handled) {
    *handled = false;
    if ((long long)B * H > 0x7FFFFFFFll) return OFD_OK;
    // rows per work unit: a multiple of the alignment group (1 / 2 / 4 rows for W % 4 == 0 / 2 / odd).  Measured on B200
    // (profiles/r1/tune_pair_groups.txt): units of ~1900-2048 pixels (1024 threads x 2 px, one CTA per SM, deep rings) reach 92-93 %
    // of the HBM peak at every width tried, 1000-1500 pixel units only 87-89 %; rows of 512-656 pixels are best left alone (two
    // CTAs per SM with the deepest rings: 96.8 % on the 480x640 headline against 93.7 % with three rows per unit).
    const int align_rows = (W % 4 == 0) ? 1 : ((W % 2 == 0) ? 2 : 4);
    if (H % align_rows != 0) return OFD_OK;  // a frame would end inside an alignment group: one-row kernel
    int G;
    if (align_rows == 1 && W >= 512 && W <= 656) {
        G = 1;
    } else {
        G = (2048 / (align_rows * W)) * align_rows;
        if (G < align_rows) G = align_rows;
    }
    if (const char* e = std::getenv("OFD_PAIR_GROUP")) {
        const int g = std::atoi(e);
        if (g > 0 && g % align_rows == 0) G = g;
    }
    if (G > H) G = (H / align_rows) * align_rows;
    const int VW = G * W;
    int in_stages = 0, out_stages = 0;
    size_t smem = 0;
    if (!choose_pair_plan<DT>(VW, &in_stages, &out_stages, &smem)) return OFD_OK;  // row too wide: fall back to the one-row kernel
    // each thread owns NITER pixels of a (virtual) row; threads = ceil(VW / NITER) rounded up to a warp
    int niter = OFD_PAIR_NITER;
    while ((VW + niter - 1) / niter > 1024) niter *= 2;
    if (niter > 4 * OFD_PAIR_NITER) return OFD_OK;
    int threads = ((VW + niter - 1) / niter + 31) / 32 * 32;
    auto kern = G == 1 ? (niter == OFD_PAIR_NITER ? pair_rows_persistent<DT, OFD_PAIR_NITER, PAIR_ROW>
                                                  : (niter == 2 * OFD_PAIR_NITER ? pair_rows_persistent<DT, 2 * OFD_PAIR_NITER, PAIR_ROW>
                                                                                 : pair_rows_persistent<DT, 4 * OFD_PAIR_NITER, PAIR_ROW>))
                       : (niter == OFD_PAIR_NITER ? pair_rows_persistent<DT, OFD_PAIR_NITER, PAIR_GROUPED>
                                                  : (niter == 2 * OFD_PAIR_NITER ? pair_rows_persistent<DT, 2 * OFD_PAIR_NITER, PAIR_GROUPED>
                                                                                 : pair_rows_persistent<DT, 4 * OFD_PAIR_NITER, PAIR_GROUPED>));
    int sms = 0, per_sm = 0;
    {
        const int rc_plan = launch_plan(fn, (const void*)kern, threads, smem, &sms, &per_sm);
        if (rc_plan) return rc_plan;
    }
    long long grid = (long long)sms * per_sm;
    const long long units = (long long)B * ((H + G - 1) / G);
    if (grid > units) grid = units;
    if (std::getenv("OFD_DEBUG"))
        fprintf(stderr, "[ofd] %s: persistent pair kernel: G=%d rows/unit in_stages=%d out_stages=%d niter=%d threads=%d smem=%zu B, %d CTAs/SM x %d SMs -> grid %lld\n",
                fn, G, in_stages, out_stages, niter, threads, smem, per_sm, sms, grid);
    kern<<<(unsigned)grid, threads, smem, st>>>(img0, depth0, sBf, img1, depth1, back_flow, flow, valid, collision, counters, B, H, W, G,
                                                in_stages, out_stages, PairNoTable{});
    *handled = true;
    return check_launch(fn);
}

// One launch for a ragged batch (up to kPairRaggedMax frames): every frame is cut into units of G_i rows with G_i * W_i <= the
// batch's unit width, and the persistent CTAs walk the concatenated unit list.  *handled stays false when a frame does not fit
// the TMA alignment rules (the caller then launches frame by frame).
template <typename DT>
static int launch_pair_ragged(const char* fn, const float* img0, const DT* depth0, const float* sBf, int n, const int* Hs,
                              const int* Ws, const size_t* offs, float* img1, float* depth1, float* back_flow, float* flow,
                              float* valid, float* collision, uint64_t* counters, cudaStream_t st, bool* handled, bool any = false) {
    *handled = false;
    PairRaggedTable tab;
    // units may start at any pixel (shifted shared-memory rows, scalar head / tail stores), so rows need no alignment groups.
    // any == false: the planes themselves start on 16-byte boundaries - every H*W and offset is a multiple of 4 pixels;
    // any == true (PAIR_RAGGED_ANY): no such rule, but the caller guarantees that the (at most 3) elements after every frame exist
    for (int i = 0; i < n && !any; ++i)
        if (((size_t)Hs[i] * Ws[i]) % 4 != 0 || offs[i] % 4 != 0) return OFD_OK;
    int vwmax = 2048;  // pixels per unit (1024 threads x 2); wider rows travel alone
    if (const char* e = std::getenv("OFD_PAIR_RAGGED_UNIT")) {  // tuning knob
        const int v = std::atoi(e);
        if (v >= 64 && v <= 2880) vwmax = v;
    }
    long long units = 0;
    int used = 0;  // widest unit actually formed
    for (int i = 0; i < n; ++i) {
        const int W = Ws[i], H = Hs[i];
        int G = vwmax / W;
        if (G < 1) G = 1;
        if (G > H) G = H;
        tab.img[i].H = H, tab.img[i].W = W, tab.img[i].G = G, tab.img[i].unit0 = (int)units;
        tab.img[i].off = (unsigned long long)offs[i];
        units += (H + G - 1) / G;
        if (G * W > used) used = G * W;
        if (units > 0x7FFFFFFFll) return OFD_OK;
    }
    tab.n = n, tab.total_units = (int)units;
    tab.img[n].H = tab.img[n].W = tab.img[n].G = 1, tab.img[n].unit0 = (int)units, tab.img[n].off = 0;
    const int VW = ((used + 3) & ~3) + 8;  // shared-memory row stride: room for the alignment shift and the rounded-up copy size
    int in_stages = 0, out_stages = 0;
    size_t smem = 0;
    if (!choose_pair_plan<DT>(VW, &in_stages, &out_stages, &smem)) return OFD_OK;
    int niter = OFD_PAIR_NITER;
    while ((used + niter - 1) / niter > 1024) niter *= 2;
    if (niter > 4 * OFD_PAIR_NITER) return OFD_OK;
    int threads = ((used + niter - 1) / niter + 31) / 32 * 32;
    if (threads < 96) threads = 96;  // up to 60 threads (of the first three warps) write the unaligned head / tail pixels
    auto kern = niter == OFD_PAIR_NITER ? pair_rows_persistent<DT, OFD_PAIR_NITER, PAIR_RAGGED>
                                        : (niter == 2 * OFD_PAIR_NITER ? pair_rows_persistent<DT, 2 * OFD_PAIR_NITER, PAIR_RAGGED>
                                                                       : pair_rows_persistent<DT, 4 * OFD_PAIR_NITER, PAIR_RAGGED>);
    if (any)
        kern = niter == OFD_PAIR_NITER ? pair_rows_persistent<DT, OFD_PAIR_NITER, PAIR_RAGGED_ANY>
                                       : (niter == 2 * OFD_PAIR_NITER ? pair_rows_persistent<DT, 2 * OFD_PAIR_NITER, PAIR_RAGGED_ANY>
                                                                      : pair_rows_persistent<DT, 4 * OFD_PAIR_NITER, PAIR_RAGGED_ANY>);
    int sms = 0, per_sm = 0;
    {
        const int rc_plan = launch_plan(fn, (const void*)kern, threads, smem, &sms, &per_sm);
        if (rc_plan) return rc_plan;
    }
    long long grid = (long long)sms * per_sm;
    if (grid > units) grid = units;
    if (std::getenv("OFD_DEBUG"))
        fprintf(stderr, "[ofd] %s: ragged persistent pair kernel: %d frames%s, %lld units of <= %d px, in_stages=%d out_stages=%d niter=%d threads=%d smem=%zu B, %d CTAs/SM -> grid %lld\n",
                fn, n, any ? " (per-plane phases)" : "", units, VW, in_stages, out_stages, niter, threads, smem, per_sm, grid);
    kern<<<(unsigned)grid, threads, smem, st>>>(img0, depth0, sBf, img1, depth1, back_flow, flow, valid, collision, counters, n, 0, VW, 1,
                                                in_stages, out_stages, tab);
    *handled = true;
    return check_launch(fn);
}


template <typename DT>
static int launch_pair(const char* fn, const float* img0, const DT* depth0, const float* sBf, int B, int H, int W,
                       float* img1, float* depth1, float* back_flow, float* flow, float* valid, float* collision,
                       uint64_t* counters, cudaStream_t st) {
    // TMA needs 16-byte addresses and sizes: plane bases are aligned when the buffers are and H*W % 4 == 0 (checked through H %
    // align_rows in launch_pair_persistent); rows that are not a multiple of 4 pixels travel in groups of 2 or 4 rows
    const bool aligned = (((uintptr_t)img0 | (uintptr_t)depth0 | (uintptr_t)img1 | (uintptr_t)depth1 |
                           (uintptr_t)back_flow | (uintptr_t)flow | (uintptr_t)valid | (uintptr_t)collision) % 16 == 0);
    int done = 0;  // frames already synthesised (per-plane ragged kernel); the rest goes to the one-row kernel
    if (aligned) {
        bool handled = false;
        int rc = OFD_OK;
        // odd widths would need 4-row alignment groups (units of 4 W pixels: 74 % of the HBM peak at 480x641); the shift-capable
        // ragged kernel cuts them into ~2000-pixel units that start on any pixel (84 %) - tried first when the planes allow it
        const bool ragged_first = (W % 2 == 1) && (((size_t)H * W) % 4 == 0) && std::getenv("OFD_PAIR_GROUP") == nullptr;
        if (!ragged_first) {
            rc = launch_pair_persistent<DT>(fn, img0, depth0, sBf, B, H, W, img1, depth1, back_flow, flow, valid, collision,
                                            counters, st, &handled);
            if (rc || handled) return rc;
        }
        // also: rows too wide for an alignment group (W % 4 != 0 and 2 or 4 rows do not fit in shared memory) travel one row per
        // unit through the same kernel, as a batch of B equal frames
        if (((size_t)H * W) % 4 == 0) {
            const size_t hw4 = (size_t)H * W;
            for (int b0 = 0; b0 < B; b0 += kPairRaggedMax) {
                const int n = (B - b0) < kPairRaggedMax ? (B - b0) : kPairRaggedMax;
                int Hs[kPairRaggedMax], Ws[kPairRaggedMax];
                size_t offs[kPairRaggedMax];
                for (int i = 0; i < n; ++i) Hs[i] = H, Ws[i] = W, offs[i] = (size_t)(b0 + i) * hw4;
                rc = launch_pair_ragged<DT>(fn, img0, depth0, sBf + b0, n, Hs, Ws, offs, img1, depth1, back_flow, flow, valid, collision,
                                            counters, st, &handled);
                if (rc) return rc;
                if (!handled) break;  // (only possible on the first chunk: every chunk has the same shape)
            }
            if (handled) return OFD_OK;
        } else {
            // H*W not a multiple of 4 (e.g. 375x1242): the planes of a frame start on different 16-byte phases - the per-plane form
            // of the ragged kernel.  Its loads read up to 3 elements past a frame, so the last frame is left to the one-row kernel
            // when the batch does not end on a 16-byte boundary.
            const size_t hw1 = (size_t)H * W;
            const int Bp = ((size_t)B * hw1) % 4 == 0 ? B : B - 1;
            for (int b0 = 0; b0 < Bp; b0 += kPairRaggedMax) {
                const int n = (Bp - b0) < kPairRaggedMax ? (Bp - b0) : kPairRaggedMax;
                int Hs[kPairRaggedMax], Ws[kPairRaggedMax];
                size_t offs[kPairRaggedMax];
                for (int i = 0; i < n; ++i) Hs[i] = H, Ws[i] = W, offs[i] = (size_t)(b0 + i) * hw1;
                rc = launch_pair_ragged<DT>(fn, img0, depth0, sBf + b0, n, Hs, Ws, offs, img1, depth1, back_flow, flow, valid, collision,
                                            counters, st, &handled, true);
                if (rc) return rc;
                if (!handled) break;
                done = b0 + n;
            }
            if (done == B) return OFD_OK;
        }
        if (ragged_first) {
            rc = launch_pair_persistent<DT>(fn, img0, depth0, sBf, B, H, W, img1, depth1, back_flow, flow, valid, collision,
                                            counters, st, &handled);
            if (rc || handled) return rc;
        }
    }
    const size_t smem = PairSmem<DT>::bytes(W);
    if (smem > 227 * 1024) return fail(OFD_E_SHAPE, "%s: W=%d needs %zu B of shared memory per row (max 227 KB)", fn, W, smem);
    const bool bulk = (W % 4 == 0) && (((uintptr_t)img0 | (uintptr_t)depth0) % 16 == 0);
    int threads = W >= 512 ? 256 : (W >= 128 ? 128 : 64);
    auto kern = bulk ? pair_row_kernel<DT, true> : pair_row_kernel<DT, false>;
    {
        const int rc_smem = ensure_dynamic_smem(fn, (const void*)kern, smem);
        if (rc_smem) return rc_smem;
    }
    const size_t hw = (size_t)H * W;
    for (int b0 = done; b0 < B; b0 += 65535) {
        const int Bc = (B - b0) < 65535 ? (B - b0) : 65535;
        dim3 grid(H, Bc);
        kern<<<grid, threads, smem, st>>>(img0 + (size_t)b0 * 3 * hw, depth0 + (size_t)b0 * hw, sBf + b0,
                                          img1 + (size_t)b0 * 3 * hw, depth1 + (size_t)b0 * hw,
                                          back_flow + (size_t)b0 * 2 * hw, flow ? flow + (size_t)b0 * 2 * hw : nullptr,
                                          valid + (size_t)b0 * hw, collision ? collision + (size_t)b0 * hw : nullptr,
                                          counters, H, W);
        int rc = check_launch(fn);
        if (rc) return rc;
    }
    return OFD_OK;
}

// Convert.depth_to_disparity + disparity_to_flow as a stand-alone elementwise op (preprocess.py:239-254)
template <typename DT>
__global__ void __launch_bounds__(256) disparity_flow_kernel(const DT* __restrict__ depth, const float* __restrict__ sBf,
                                                            DT* __restrict__ flow, size_t hw) {
    const int b = blockIdx.y;
    const DT s = (DT)sBf[b];
    const DT* d = depth + (size_t)b * hw;
    DT* f = flow + (size_t)b * 2 * hw;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (size_t)gridDim.x * blockDim.x) {
        const DT disp = s / d[p];
        f[p] = disp * (DT)-1.0;
        f[hw + p] = (DT)-0.0;
    }
}

}  // namespace ofd

using namespace ofd;

template <typename DT>
static int pair_ragged(const char* fn, const float* img0, const DT* depth0, const float* sBf, int n_images, const int* Hs,
                       const int* Ws, const size_t* offs, float* img1, float* depth1, float* back_flow, float* flow, float* valid,
                       float* collision, uint64_t* counters, cudaStream_t st) {
    const bool aligned = (((uintptr_t)img0 | (uintptr_t)depth0 | (uintptr_t)img1 | (uintptr_t)depth1 | (uintptr_t)back_flow |
                           (uintptr_t)flow | (uintptr_t)valid | (uintptr_t)collision) % 16 == 0);
    auto one_frame = [&](int i) {  // one launch for frame i alone (whatever kernel its shape allows)
        const size_t o = offs[i];
        return launch_pair<DT>(fn, img0 + 3 * o, depth0 + o, sBf + i, 1, Hs[i], Ws[i], img1 + 3 * o, depth1 + o, back_flow + 2 * o,
                               flow ? flow + 2 * o : nullptr, valid + o, collision ? collision + o : nullptr, counters, st);
    };
    // frames [lo, hi) in launches of up to kPairRaggedMax; sBf is indexed by the frame's position in the launch, so each launch gets
    // its slice of sBf while the offsets stay batch-wide
    auto range = [&](int lo, int hi, bool any) {
        for (int i0 = lo; i0 < hi; i0 += kPairRaggedMax) {
            const int n = (hi - i0) < kPairRaggedMax ? (hi - i0) : kPairRaggedMax;
            bool handled = false;
            if (aligned) {
                int rc = launch_pair_ragged<DT>(fn, img0, depth0, sBf + i0, n, Hs + i0, Ws + i0, offs + i0, img1, depth1, back_flow, flow,
                                                valid, collision, counters, st, &handled, any);
                if (rc) return rc;
            }
            if (handled) continue;
            for (int i = i0; i < i0 + n; ++i) {  // a row too wide for shared memory, unaligned buffers: one launch per frame
                int rc = one_frame(i);
                if (rc) return rc;
            }
        }
        return (int)OFD_OK;
    };
    bool uniform = true;  // every plane of every frame starts on a 16-byte boundary
    int last = 0;         // the frame that ends the buffers
    for (int i = 0; i < n_images; ++i) {
        const size_t hw = (size_t)Hs[i] * Ws[i];
        if (hw % 4 != 0 || offs[i] % 4 != 0) uniform = false;
        if (offs[i] + hw > offs[last] + (size_t)Hs[last] * Ws[last]) last = i;
    }
    if (uniform) return range(0, n_images, false);
    // per-plane phases: the aligned-superset loads read up to 3 elements past a frame; they exist for every frame but the one that
    // ends the buffers (whose sizes are not known here) - it travels alone unless it ends on a 16-byte boundary
    const bool last_alone = (offs[last] + (size_t)Hs[last] * Ws[last]) % 4 != 0;
    if (!last_alone) return range(0, n_images, true);
    int rc = range(0, last, true);
    if (rc == OFD_OK) rc = range(last + 1, n_images, true);
    if (rc == OFD_OK) rc = one_frame(last);
    return rc;
}

extern "C" {

int ofd_disparity_pair(const float* img0, const void* depth0, int depth_dtype, const float* sBf, int B, int H, int W,
                       float* img1, float* depth1, float* back_flow, float* flow, float* valid, float* collision,
                       uint64_t* counters, ofd_stream_t stream) {
    const char* fn = "ofd_disparity_pair";
    if (depth_dtype != OFD_F32 && depth_dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad depth dtype %d", fn, depth_dtype);
    if (B < 0 || H < 0 || W < 0) return fail(OFD_E_SHAPE, "%s: negative dimension", fn);
    if ((size_t)H * (size_t)W >= ((size_t)1 << 31)) return fail(OFD_E_SHAPE, "%s: H*W must be < 2^31", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!img0 || !depth0 || !sBf || !img1 || !depth1 || !back_flow || !valid)
        return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    cudaStream_t st = (cudaStream_t)stream;
    if (depth_dtype == OFD_F32)
        return launch_pair<float>(fn, img0, (const float*)depth0, sBf, B, H, W, img1, depth1, back_flow, flow, valid,
                                  collision, counters, st);
    return launch_pair<double>(fn, img0, (const double*)depth0, sBf, B, H, W, img1, depth1, back_flow, flow, valid,
                               collision, counters, st);
}

int ofd_disparity_pair_ragged(const float* img0, const void* depth0, int depth_dtype, const float* sBf, int n_images,
                              const int* H_host, const int* W_host, const size_t* offset_host, float* img1, float* depth1,
                              float* back_flow, float* flow, float* valid, float* collision, uint64_t* counters,
                              ofd_stream_t stream) {
    const char* fn = "ofd_disparity_pair_ragged";
    if (depth_dtype != OFD_F32 && depth_dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad depth dtype %d", fn, depth_dtype);
    if (n_images < 0) return fail(OFD_E_SHAPE, "%s: negative n_images", fn);
    if (n_images == 0) return OFD_OK;
    if (!H_host || !W_host || !offset_host) return fail(OFD_E_NULL, "%s: NULL shape table", fn);
    if (!img0 || !depth0 || !sBf || !img1 || !depth1 || !back_flow || !valid)
        return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    for (int i = 0; i < n_images; ++i) {
        if (H_host[i] <= 0 || W_host[i] <= 0) return fail(OFD_E_SHAPE, "%s: image %d has a non-positive dimension", fn, i);
        if ((size_t)H_host[i] * (size_t)W_host[i] >= ((size_t)1 << 31)) return fail(OFD_E_SHAPE, "%s: image %d: H*W must be < 2^31", fn, i);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (depth_dtype == OFD_F32)
        return pair_ragged<float>(fn, img0, (const float*)depth0, sBf, n_images, H_host, W_host, offset_host, img1, depth1,
                                  back_flow, flow, valid, collision, counters, st);
    return pair_ragged<double>(fn, img0, (const double*)depth0, sBf, n_images, H_host, W_host, offset_host, img1, depth1,
                               back_flow, flow, valid, collision, counters, st);
}

int ofd_disparity_flow(const void* depth, int depth_dtype, const float* sBf, int B, int H, int W, void* flow,
                       ofd_stream_t stream) {
    const char* fn = "ofd_disparity_flow";
    if (depth_dtype != OFD_F32 && depth_dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad depth dtype %d", fn, depth_dtype);
    if (B < 0 || H < 0 || W < 0 || B > 65535) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!depth || !sBf || !flow) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    const size_t hw = (size_t)H * W;
    dim3 grid((unsigned)((hw + 1023) / 1024), B);
    if (depth_dtype == OFD_F32)
        disparity_flow_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)depth, sBf, (float*)flow, hw);
    else
        disparity_flow_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)depth, sBf, (double*)flow, hw);
    return check_launch(fn);
}

}  // extern "C"
