// ofd_inpaint.cu — Telea fast-marching inpainting on the GPU: the fill of utils.inpaint (utils.py:136-151), which the reference
// hands to OpenCV on the host (cv2.inpaint(img_u8, mask, 3, cv2.INPAINT_TELEA)) 95 times per frame.
//
// OpenCV's algorithm (the published one: A. Telea, "An image inpainting technique based on the fast marching method", 2004, as
// implemented in opencv/modules/photo/src/inpaint.cpp, which is NOT part of /root/reference - third-party, unpinned, restated from
// its published form):
//   1. the hole's 4-neighbour ring ("band") gets T = 0; a fast-marching pass over the known pixels within `range` of the hole gives
//      them T = -distance; hole pixels get T = +distance when they are reached;
//   2. hole pixels are filled in order of increasing T (a heap): a pixel is computed when the first of its 4-neighbours is popped -
//      its T from the upwind quadrant solve, its colour as the weighted mean of the already known pixels within `range`
//      (weights: direction x distance x level-set factors) plus a gradient term - and is then known to every later pixel.
// The heap makes this inherently serial.  Here the march advances in LAYERS: layer k holds the hole pixels that have a 4-neighbour in
// layers < k (layer 0 = known).  All pixels of a layer are computed in parallel from pixels of earlier layers only, with the same
// per-pixel arithmetic, in the same operand order, as OpenCV's.  For thin disocclusion bands the layer order and the heap order give
// nearly the same result; deep holes differ more (the fill order inside a ring of equal depth differs).  Bit equality with the
// heap order is therefore NOT claimed: tests/test_gpu_parity.py measures the fraction of bytes that differ from cv2.inpaint by more
// than one grey level and states the bound; synthesis.inpaint keeps backend="cv2" selectable.
//
// One cooperative kernel per batch: init -> outward ring march (2 * range layers) -> inward layers until no hole pixel is left,
// each layer in two phases (compute into a staging list; commit + enqueue the next frontier) separated by grid-wide barriers, so
// that every read of a layer sees the state at the layer's start (deterministic results).
#include <cooperative_groups.h>

#include <cstdlib>

#include "ofd_common.cuh"

namespace cg = cooperative_groups;

namespace ofd {

constexpr float T_FAR = 1.0e6f;         // inpaint.cpp: cvSet(t, 1.0e6f)
constexpr unsigned short L_INSIDE = 0xFFFFu;

struct TeleaParams {
    const float* img;          // [B,3,H,W] float32, uint8-valued
    const unsigned char* mask; // [B,1,H,W], != 0 = fill
    float* out;                // [B,3,H,W]
    int B, H, W, range;
    // workspace (per batch)
    unsigned char* u8;         // [B,3,H,W] working image
    unsigned char* u8o;        // [B,3,H,W] the original bytes (read where OpenCV's border index shifts hit a cell that is still unknown)
    unsigned short* L;         // [B,(H+2),(W+2)] layer: 0 known, 0xFFFF unfilled hole, k filled in layer k
    float* T;                  // [B,(H+2),(W+2)]
    unsigned char* G;          // [B,(H+2),(W+2)] outward march: 0 other, 1 band, 255 ring (uncomputed), k+1 computed in out-layer k
    unsigned int* Q;           // [B,(H+2),(W+2)] 1 = already enqueued
    unsigned int* list[3];     // rotating frontier lists: extended pixel index of each entry
    unsigned int* listb[3];    // frame index of each entry
    float* stageT;             // ring march: staged T per extended pixel
    unsigned int* count;       // [0..2] list sizes, [3] = layers done, [4] = filled pixels
};

// inpaint.cpp FastMarching_solve: upwind quadrant solve from the two neighbours (i1,j1), (i2,j2); `known(q)` = f != INSIDE
__device__ __forceinline__ float fmm_solve(float a11, bool k1, float a22, bool k2) {
    double sol;
    const double m12 = a11 < a22 ? (double)a11 : (double)a22;
    if (k1) {
        if (k2) {
            const double d = (double)a11 - (double)a22;
            if (fabs(d) >= 1.0)
                sol = 1 + m12;
            else
                sol = ((double)a11 + (double)a22 + sqrt(2 - d * d)) * 0.5;
        } else {
            sol = 1 + (double)a11;
        }
    } else if (k2) {
        sol = 1 + (double)a22;
    } else {
        sol = 1 + m12;
    }
    return (float)sol;
}

__device__ __forceinline__ float min4(float a, float b, float c, float d) { return fminf(fminf(a, b), fminf(c, d)); }

#ifndef OFD_TELEA_MINB
#define OFD_TELEA_MINB 4  // 64 registers, 32 warps per SM: batch of 90 fills 14.9 -> 13.3 ms, batch of 9 2.80 -> 2.67 ms (profiles/r2/tune_telea_minb.txt)
#endif
__global__ void __launch_bounds__(256, OFD_TELEA_MINB) telea_kernel(const __grid_constant__ TeleaParams P) {
    cg::grid_group grid = cg::this_grid();
    const int H = P.H, W = P.W, EW = W + 2, EH = H + 2, range = P.range;
    const size_t hw = (size_t)H * W, ehw = (size_t)EH * EW;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    // per-warp tap table of phase 2: (2 range + 1)^2 taps x TAP_F floats {w, w*I[3], w*gix*rx[3], w*giy*ry[3], used}
    constexpr int TAP_F = 11;
    extern __shared__ float s_dyn[];
    // the taps of a pixel: the offsets (dk, dl) inside the circle dk^2 + dl^2 <= range^2 WITHOUT the centre (always a hole pixel), in
    // OpenCV's loop order (k outer, l inner): 28 of the 49 window cells at range 3, so one round of the warp's lanes covers them all
    const int D = 2 * range + 1;
    int NT = 0;
    for (int t = 0; t < D * D; ++t) {
        const int dk = t / D - range, dl = t - (t / D) * D - range;
        NT += (dk * dk + dl * dl <= range * range) && (dk | dl) != 0;
    }
    const int lane = threadIdx.x & 31;
    const size_t gwarp = tid >> 5, nwarps = nth >> 5;
    const int NTP = (D * D + 3) & ~3;                                 // table stride (>= NT, the host's allocation unit)
    float* s_dst = s_dyn;                                              // [NT]  distance factor of every tap offset
    int* s_off = reinterpret_cast<int*>(s_dyn + NTP);                 // [NT]  (dk << 16) | (dl & 0xFFFF)
    float* tw = s_dyn + 2 * NTP + (size_t)(threadIdx.x >> 5) * NT * TAP_F;
    if (threadIdx.x == 0) {
        int n = 0;
        for (int t = 0; t < D * D; ++t) {
            const int dk = t / D - range, dl = t - (t / D) * D - range;
            if (dk * dk + dl * dl > range * range || (dk | dl) == 0) continue;
            const float ry = (float)(-dk), rx = (float)(-dl);
            const float len2 = rx * rx + ry * ry;
            s_dst[n] = (float)(1. / ((double)len2 * sqrt((double)len2)));  // inpaint.cpp: 1 / (|r|^2 * sqrt(|r|^2))
            s_off[n] = (dk << 16) | (dl & 0xFFFF);
            ++n;
        }
    }
    __syncthreads();

    // ---- phase 0: working image, flags, T, band, ring ------------------------------------------------------------------------------
    for (size_t e = tid; e < (size_t)P.B * 3 * hw; e += nth) {
        float v = P.img[e];
        v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
        const unsigned char u = (unsigned char)(int)v;  // .astype(np.uint8): truncation (utils.py:147)
        P.u8[e] = u;
        P.u8o[e] = u;
    }
    if (tid < 8) P.count[tid] = 0;
    for (size_t e = tid; e < (size_t)P.B * ehw; e += nth) {
        const int b = (int)(e / ehw);
        const int p = (int)(e - (size_t)b * ehw);
        const int i = p / EW, j = p - i * EW;  // extended coordinates: interior is 1..H x 1..W
        const unsigned char* m = P.mask + (size_t)b * hw;
        auto hole = [&](int ii, int jj) -> bool { return ii >= 1 && ii <= H && jj >= 1 && jj <= W && m[(size_t)(ii - 1) * W + (jj - 1)] != 0; };
        const bool interior = i >= 1 && i <= H && j >= 1 && j <= W;
        const bool me = interior && hole(i, j);
        bool band = false, ring = false;
        if (interior && !me) {
            band = hole(i - 1, j) || hole(i + 1, j) || hole(i, j - 1) || hole(i, j + 1);  // 3x3 cross dilation minus the mask
            if (!band) {
                for (int di = -range; di <= range && !ring; ++di)
                    for (int dj = -range; dj <= range; ++dj)
                        if (hole(i + di, j + dj)) {
                            ring = true;  // (2 range + 1)^2 rectangle dilation minus mask minus band
                            break;
                        }
            }
        }
        P.L[e] = me ? L_INSIDE : 0;
        P.T[e] = band ? 0.0f : T_FAR;
        P.G[e] = band ? 1 : (ring ? 255 : 0);
        P.Q[e] = 0;
    }
    grid.sync();

    // ---- phase 1: outward march over the ring (icvCalcFMM(out, t, Out, negate = true)) -------------------------------------------------
    // out-layer k: ring pixels with a 4-neighbour popped earlier (band = layer 0).  In that pass every non-ring pixel counts as known
    // with its current T (band 0, everything else 1e6), exactly as the flags of OpenCV's `out` matrix say.
    for (int k = 1; k <= 2 * range; ++k) {
        for (size_t e = tid; e < (size_t)P.B * ehw; e += nth) {
            if (P.G[e] != 255) continue;
            const int b = (int)(e / ehw);
            const int p = (int)(e - (size_t)b * ehw);
            const unsigned char* G = P.G + (size_t)b * ehw;
            const float* T = P.T + (size_t)b * ehw;
            auto popped = [&](int q) -> bool { const unsigned char g = G[q]; return g >= 1 && g <= k; };
            if (!(popped(p - EW) || popped(p + EW) || popped(p - 1) || popped(p + 1))) continue;
            auto known = [&](int q) -> bool { return G[q] <= k; };  // 0 (not ring), band, or computed in an earlier out-layer
            const int up = p - EW, dn = p + EW, lf = p - 1, rt = p + 1;
            const float d = min4(fmm_solve(T[up], known(up), T[lf], known(lf)), fmm_solve(T[dn], known(dn), T[lf], known(lf)),
                                 fmm_solve(T[up], known(up), T[rt], known(rt)), fmm_solve(T[dn], known(dn), T[rt], known(rt)));
            P.stageT[e] = d;  // staged: committed after the barrier (reads of this layer see the layer's start state)
            P.Q[e] = 2;
        }
        grid.sync();
        for (size_t e = tid; e < (size_t)P.B * ehw; e += nth)
            if (P.Q[e] == 2) {
                P.T[e] = P.stageT[e];
                P.G[e] = (unsigned char)(k + 1);
                P.Q[e] = 0;
            }
        grid.sync();
    }
    // negate the ring distances; build the first frontier: hole pixels with an interior known 4-neighbour
    for (size_t e = tid; e < (size_t)P.B * ehw; e += nth) {
        const unsigned char g = P.G[e];
        if (g >= 2 && g != 255) P.T[e] = -P.T[e];
        if (P.L[e] == L_INSIDE) {
            const int b = (int)(e / ehw);
            const int p = (int)(e - (size_t)b * ehw);
            const unsigned short* L = P.L + (size_t)b * ehw;
            auto trig = [&](int q) -> bool {
                const int qi = q / EW, qj = q - qi * EW;
                return L[q] == 0 && qi >= 1 && qi <= H && qj >= 1 && qj <= W;
            };
            if (trig(p - EW) || trig(p + EW) || trig(p - 1) || trig(p + 1)) {
                const unsigned int slot = atomicAdd(&P.count[0], 1u);
                P.list[0][slot] = (unsigned int)p;
                P.listb[0][slot] = (unsigned int)b;
                P.Q[e] = 1;
            }
        }
    }
    grid.sync();

    // ---- phase 2: inward layers (icvTeleaInpaintFMM) -------------------------------------------------------------------------------------
    // ONE grid-wide barrier per layer.  A layer's pixels are computed from the state at the layer's start and committed at once (T, colour,
    // L = layer, next frontier): that is race-free because a reader never consumes a cell of the current layer - it tests L[q] < layer
    // first (a cell being committed reads as 0xFFFF or as `layer`, unknown either way), takes T = 1e6 for unknown cells (what they hold at
    // the layer's start) and, where OpenCV's border index shifts make it read the IMAGE of a cell that is still unknown, takes the cell's
    // original byte from `u8o` (again what `u8` holds at the layer's start).  Three frontier lists rotate: read k % 3, append (k+1) % 3, and
    // the counter of (k+2) % 3 - the next layer's append target, untouched during this layer - is zeroed.
    for (unsigned int layer = 1; layer < L_INSIDE; ++layer) {
        const int cur = (int)((layer - 1) % 3u), nxt = (int)(layer % 3u), zro = (int)((layer + 1) % 3u);
        const unsigned int n = P.count[cur];
        if (n == 0) break;
        if (tid == 0) P.count[zro] = 0, P.count[3] = layer, P.count[4] += n;
        // one WARP per pixel: the lanes evaluate the (2 range + 1)^2 taps in parallel (flags, weights, gradients: the expensive part) into a
        // per-warp shared-memory table; lanes 0-2 (one per colour) then add the tap contributions in OpenCV's tap order (k outer, l inner),
        // so every float32 sum is formed in the same order as the serial code - the result does not depend on the lane layout.  (A thread per
        // pixel ran a layer of a few thousand pixels at ~1 warp per SM: 130 us per layer; profiles/r2/inpaint_report_v1_thread_per_pixel.json.)
        for (size_t e = gwarp; e < n; e += nwarps) {
            const int b = (int)P.listb[cur][e];
            const int p = (int)P.list[cur][e];
            const int i = p / EW, j = p - i * EW;
            const size_t eb = (size_t)b * ehw;
            unsigned short* L = P.L + eb;
            float* T = P.T + eb;
            unsigned char* I = P.u8 + (size_t)b * 3 * hw;
            const unsigned char* Io = P.u8o + (size_t)b * 3 * hw;
            auto known = [&](int q) -> bool { return L[q] < layer; };  // f != INSIDE
            auto Tk = [&](int q) -> float { return known(q) ? T[q] : T_FAR; };
            const int up = p - EW, dn = p + EW, lf = p - 1, rt = p + 1;
            const bool k_up = known(up), k_dn = known(dn), k_lf = known(lf), k_rt = known(rt);
            const float t_up = k_up ? T[up] : T_FAR, t_dn = k_dn ? T[dn] : T_FAR, t_lf = k_lf ? T[lf] : T_FAR, t_rt = k_rt ? T[rt] : T_FAR;
            const float dist = min4(fmm_solve(t_up, k_up, t_lf, k_lf), fmm_solve(t_dn, k_dn, t_lf, k_lf), fmm_solve(t_up, k_up, t_rt, k_rt),
                                    fmm_solve(t_dn, k_dn, t_rt, k_rt));
            // gradT (t(i,j) is the value just computed)
            float gx, gy;
            if (k_rt)
                gx = k_lf ? (t_rt - t_lf) * 0.5f : (t_rt - dist);
            else
                gx = k_lf ? (dist - t_lf) : 0.0f;
            if (k_dn)
                gy = k_up ? (t_dn - t_up) * 0.5f : (t_dn - dist);
            else
                gy = k_up ? (dist - t_up) : 0.0f;
            for (int t = lane; t < NT; t += 32) {
                const int off = s_off[t];
                const int k = i + (off >> 16), l = j + (int)(short)(off & 0xFFFF);
                float* slot = tw + t * TAP_F;
                bool use = k > 0 && k < EH - 1 && l > 0 && l < EW - 1;
                const int q = k * EW + l;
                use = use && known(q);
                slot[10] = use ? 1.0f : 0.0f;
                if (!use) continue;
                const int km = k - 1 + (k == 1), kp = k - 1 - (k == EH - 2);
                const int lm = l - 1 + (l == 1), lp = l - 1 - (l == EW - 2);
                const float ry = (float)(i - k), rx = (float)(j - l);
                const float lev = (float)(1. / (1 + fabs((double)(Tk(q) - dist))));
                float dir = rx * gx + ry * gy;
                if (fabsf(dir) <= 0.01f) dir = 0.000001f;
                const float w = fabsf(s_dst[t] * lev * dir);
                const bool kr = known(q + 1), kl = known(q - 1), kd = known(q + EW), ku = known(q - EW);
                slot[0] = w;
                // image position (r, col) -> element offset, and whether the cell is known (else its original byte is read); the row /
                // column clamps only matter for frames with H < 2 or W < 2, where inpaint.cpp's km / lm shifts leave the image
                auto at = [&](int r, int col, bool& kn) -> size_t {
                    r = r < 0 ? 0 : (r > H - 1 ? H - 1 : r);
                    col = col < 0 ? 0 : (col > W - 1 ? W - 1 : col);
                    kn = L[(r + 1) * EW + (col + 1)] < layer;
                    return (size_t)r * W + col;
                };
                bool nA, nB, nC, nD, nE, nF, nG;
                const size_t oA = at(km, lp + 1, nA), oB = at(km, lm - 1, nB), oC = at(km, lm, nC), oD = at(km, lp, nD);
                const size_t oE = at(kp + 1, lm, nE), oF = at(km - 1, lm, nF), oG = at(kp, lm, nG);
                const size_t oZ = (size_t)(k - 1) * W + (l - 1);  // the tap's own pixel: known
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const unsigned char* Ic = I + (size_t)c * hw;
                    const unsigned char* Oc = Io + (size_t)c * hw;
                    auto px = [&](size_t o, bool kn) -> float { return (float)(kn ? Ic[o] : Oc[o]); };
                    float gix, giy;
                    if (kr)
                        gix = kl ? (px(oA, nA) - px(oB, nB)) * 2.0f : (px(oA, nA) - px(oC, nC));
                    else
                        gix = kl ? (px(oD, nD) - px(oB, nB)) : 0.0f;
                    if (kd)
                        giy = ku ? (px(oE, nE) - px(oF, nF)) * 2.0f : (px(oE, nE) - px(oC, nC));
                    else
                        giy = ku ? (px(oG, nG) - px(oF, nF)) : 0.0f;
                    slot[1 + c] = w * (float)Ic[oZ];  // the tap's own pixel; the km / lm shifts apply to the gradients only
                    slot[4 + c] = w * (gix * rx);
                    slot[7 + c] = w * (giy * ry);
                }
            }
            __syncwarp();
            // twelve lanes add up the taps, one running sum each (colour c = lane / 4; sum 0 = Ia, 1 = Jx, 2 = Jy, 3 = s), in tap order
            float acc = (lane & 3) == 3 ? 1.0e-20f : 0.0f;
            if (lane < 12) {
                const int c = lane >> 2, a = lane & 3;
                const int col = a == 0 ? 1 + c : (a == 1 ? 4 + c : (a == 2 ? 7 + c : 0));
                for (int t = 0; t < NT; ++t) {
                    const float* slot = tw + t * TAP_F;
                    if (slot[10] != 0.0f) acc = (a == 1 || a == 2) ? acc - slot[col] : acc + slot[col];
                }
            }
            const int base = (lane < 12 ? lane : 0) & ~3;
            const float Ia = __shfl_sync(0xFFFFFFFFu, acc, base), Jx = __shfl_sync(0xFFFFFFFFu, acc, base + 1);
            const float Jy = __shfl_sync(0xFFFFFFFFu, acc, base + 2), sw = __shfl_sync(0xFFFFFFFFu, acc, base + 3);
            if (lane < 12 && (lane & 3) == 0) {
                const int c = lane >> 2;
                const float sat = (float)(Ia / sw + (Jx + Jy) / (sqrtf(Jx * Jx + Jy * Jy) + 1.0e-20f) + 0.5f);
                int v = __float2int_rn(sat);  // saturate_cast<uchar>(float): round to nearest even, saturate
                if (!(sat == sat)) v = 0;
                I[(size_t)c * hw + (size_t)(i - 1) * W + (j - 1)] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));  // commit the colour
            }
            __syncwarp();
            if (lane == 0) {  // commit T and the layer, enqueue the unfilled 4-neighbours for the next layer
                T[p] = dist;
                L[p] = (unsigned short)layer;
                const int nb[4] = {up, lf, dn, rt};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int q = nb[t];
                    const int qi = q / EW, qj = q - qi * EW;
                    if (qi <= 0 || qj <= 0 || qi > EH - 1 || qj > EW - 1) continue;  // inpaint.cpp: (i<=0)||(j<=0)||(i>rows-1)||(j>cols-1)
                    if (L[q] == L_INSIDE && atomicExch(&P.Q[eb + q], 1u) == 0u) {
                        const unsigned int slot_n = atomicAdd(&P.count[nxt], 1u);
                        P.list[nxt][slot_n] = (unsigned int)q;
                        P.listb[nxt][slot_n] = (unsigned int)b;
                    }
                }
            }
            __syncwarp();
        }
        grid.sync();
    }
    // ---- result ------------------------------------------------------------------------------------------------------------------------------
    for (size_t e = tid; e < (size_t)P.B * 3 * hw; e += nth) P.out[e] = (float)P.u8[e];
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace ofd

using namespace ofd;

extern "C" {

size_t ofd_inpaint_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t hw = (size_t)H * W, ehw = (size_t)(H + 2) * (W + 2), nb = (size_t)B;
    size_t n = 256;
    n += 2 * align256(nb * 3 * hw);             // u8, u8o
    n += align256(nb * ehw * 2);                // L
    n += align256(nb * ehw * 4);                // T
    n += align256(nb * ehw);                    // G
    n += align256(nb * ehw * 4);                // Q
    n += 6 * align256(nb * hw * 4);             // list[3], listb[3]
    n += align256(nb * ehw * 4);                // stageT (indexed by extended pixel in the ring march)
    return n;
}

int ofd_inpaint_telea(const float* img, const uint8_t* mask, int B, int H, int W, int range, float* out, void* ws, size_t ws_bytes,
                      uint32_t* stats_host_or_null, ofd_stream_t stream) {
    const char* fn = "ofd_inpaint_telea";
    if (B < 0 || H < 0 || W < 0) return fail(OFD_E_SHAPE, "%s: negative dimension", fn);
    if (range < 1 || range > 5) return fail(OFD_E_ARG, "%s: range %d outside [1,5]", fn, range);
    if ((size_t)(H + 2) * (size_t)(W + 2) >= ((size_t)1 << 31)) return fail(OFD_E_SHAPE, "%s: frame too large", fn);
    if (B == 0 || H == 0 || W == 0) return OFD_OK;
    if (!img || !mask || !out) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    const size_t need = ofd_inpaint_workspace_bytes(B, H, W);
    if (!ws || ((uintptr_t)ws & 255) || ws_bytes < need)
        return fail(OFD_E_WORKSPACE, "%s: workspace needs %zu bytes, 256-byte aligned (got %zu)", fn, need, ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)H * W, ehw = (size_t)(H + 2) * (W + 2), nb = (size_t)B;
    unsigned char* w = (unsigned char*)ws;
    TeleaParams P = {};
    P.img = img, P.mask = mask, P.out = out, P.B = B, P.H = H, P.W = W, P.range = range;
    P.count = (unsigned int*)w, w += 256;
    P.u8 = w, w += align256(nb * 3 * hw);
    P.u8o = w, w += align256(nb * 3 * hw);
    P.L = (unsigned short*)w, w += align256(nb * ehw * 2);
    P.T = (float*)w, w += align256(nb * ehw * 4);
    P.G = w, w += align256(nb * ehw);
    P.Q = (unsigned int*)w, w += align256(nb * ehw * 4);
    for (int k = 0; k < 3; ++k) P.list[k] = (unsigned int*)w, w += align256(nb * hw * 4);
    for (int k = 0; k < 3; ++k) P.listb[k] = (unsigned int*)w, w += align256(nb * hw * 4);
    P.stageT = (float*)w, w += align256(nb * ehw * 4);
    int sms = 0, per_sm = 0;
    const int D2 = (2 * range + 1) * (2 * range + 1);
    int NT = 0;  // taps inside the circle, centre excluded
    for (int dk = -range; dk <= range; ++dk)
        for (int dl = -range; dl <= range; ++dl) NT += (dk * dk + dl * dl <= range * range) && (dk | dl) != 0;
    const size_t smem = ((size_t)2 * ((D2 + 3) & ~3) + (size_t)8 * NT * 11) * sizeof(float);  // dst + offset tables, 8 warps x NT taps x 11 floats
    int rc = launch_plan(fn, (const void*)telea_kernel, 256, smem, &sms, &per_sm);
    if (rc) return rc;
    int dev = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return fail(OFD_E_ARG, "%s: device does not support cooperative launches", fn);
    if (const char* env = getenv("OFD_TELEA_BLOCKS_PER_SM")) {  // tuning knob: fewer blocks make the grid-wide barriers cheaper
        const int v = atoi(env);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    long long blocks = (long long)sms * per_sm;
    const long long useful = (long long)((nb * ehw + 255) / 256);
    if (blocks > useful) blocks = useful < 1 ? 1 : useful;
    void* args[] = {(void*)&P};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)telea_kernel, dim3((unsigned)blocks), dim3(256), args, smem, st);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaLaunchCooperativeKernel: %s", fn, cudaGetErrorString(e));
    if (stats_host_or_null) {  // layers marched / pixels filled: a synchronising read, for tests and reports only
        e = cudaMemcpyAsync(stats_host_or_null, P.count + 3, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
    }
    return check_launch(fn);
}

}  // extern "C"
