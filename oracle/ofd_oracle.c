/*
 * ofd_oracle.c — CPU restatement of the reference's forward-warp path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (opticalflowfromdepth_b200/, dropin/) may import, link or call this file; it is the
 * checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: the reference has no tests, golden vectors or fixtures for this path, so
 * this oracle is pinned against (a) the reference's own Python (alt_cuda/fw.py prologue, preprocess.py Convert,
 * geometry.py, bilateral_filter.py) imported in the build container by tests/golden/make_golden.py (small cases, arrays
 * committed) and tests/golden/make_golden_fullsize.py (480x640 / 368x496 / ReDWeb / 1080p, SHA-256 digests committed), and
 * (b) the reference's own CUDA kernel compiled unmodified from /root/reference/alt_cuda into oracle/_ref/ and run on the
 * B200 by tests/test_gpu_parity.py::test_against_the_reference_kernel_itself.
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DLUT_INIT 1000.0f /* alt_cuda/fw_cuda_kernel.cu:58  ones_like(obj) * 1000. */

/*
 * alt_cuda/fw_cuda_kernel.cu:10-83, literally: one "thread" per channel c, each with its own depth LUT plane,
 * walking all source pixels in raster order; thread 0 also maintains valid / collision.
 * obj[B,C,H,W]; safe_y, safe_x, depth[B,1,H,W]; out[B,C,H,W]; valid, collision[B,1,H,W].  float32.
 * Returns 0, or -1 if a target index is out of range (undefined behaviour in the reference).
 */
int oracle_splat_literal(const float* obj, const float* safe_y, const float* safe_x, const float* depth, int B, int C,
                         int H, int W, float* out, float* valid, float* collision) {
    const size_t hw = (size_t)H * W;
    float* dlut = (float*)malloc(sizeof(float) * (size_t)B * C * hw);
    if (!dlut && B && C && hw) return -2;
    for (size_t k = 0; k < (size_t)B * C * hw; ++k) { /* :57-58 */
        out[k] = 0.0f;
        dlut[k] = DLUT_INIT;
    }
    for (size_t k = 0; k < (size_t)B * hw; ++k) valid[k] = 0.0f, collision[k] = 0.0f; /* :59-60 */
    int rc = 0;
    for (int b = 0; b < B; ++b) {       /* blockIdx.x  :25 */
        for (int c = 0; c < C; ++c) {   /* threadIdx.x :26 */
            const float* o = obj + ((size_t)b * C + c) * hw;
            float* op = out + ((size_t)b * C + c) * hw;
            float* dl = dlut + ((size_t)b * C + c) * hw;
            for (int j = 0; j < H; ++j) {      /* :28 */
                for (int i = 0; i < W; ++i) {  /* :29 */
                    const size_t p = (size_t)b * hw + (size_t)j * W + i;
                    const float xf = safe_x[p], yf = safe_y[p]; /* :31-32 */
                    if (!(xf > -1.0f && xf < (float)W && yf > -1.0f && yf < (float)H)) {
                        rc = -1; /* reference: out-of-bounds write */
                        continue;
                    }
                    const int x = (int)xf, y = (int)yf; /* accessor index: float -> int truncation */
                    const size_t t = (size_t)y * W + x;
                    if (depth[p] < dl[t]) { /* :34 strict '<' */
                        op[t] = o[(size_t)j * W + i]; /* :35 */
                        dl[t] = depth[p];             /* :36 */
                    }
                    if (c == 0) { /* :38-45 */
                        valid[(size_t)b * hw + t] = 1.0f;
                        collision[(size_t)b * hw + t] = (dl[t] != 1000.) ? 0.0f : 1.0f;
                    }
                }
            }
        }
    }
    free(dlut);
    return rc;
}

/*
 * alt_cuda/fw.py:27-42: p1 = p0 + flow (p0 float32 grid; promoted to the flow dtype), clamp to the image, truncate
 * via int64, back to float32.  flow[2,H,W] -> safe_x, safe_y[H,W] (float32).  A NaN coordinate is reported as -1e30
 * (the reference's NaN -> int64 cast is undefined; the product drops such sources).
 */
void oracle_fw_targets_f32(const float* flow, int H, int W, float* safe_x, float* safe_y) {
    const size_t hw = (size_t)H * W;
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            const size_t p = (size_t)j * W + i;
            float px = (float)i + flow[p], py = (float)j + flow[hw + p]; /* :31 */
            if (px != px || py != py) {
                safe_x[p] = safe_y[p] = -1e30f;
                continue;
            }
            px = px < 0.0f ? 0.0f : px; /* :38 clamp(min=0, max=w-1) */
            px = px > (float)(W - 1) ? (float)(W - 1) : px;
            py = py < 0.0f ? 0.0f : py; /* :37 */
            py = py > (float)(H - 1) ? (float)(H - 1) : py;
            safe_x[p] = (float)(int64_t)px; /* :42 .type(int64).type(float32) */
            safe_y[p] = (float)(int64_t)py; /* :41 */
        }
}

void oracle_fw_targets_f64(const double* flow, int H, int W, float* safe_x, float* safe_y) {
    const size_t hw = (size_t)H * W;
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            const size_t p = (size_t)j * W + i;
            double px = (double)(float)i + flow[p], py = (double)(float)j + flow[hw + p];
            if (px != px || py != py) {
                safe_x[p] = safe_y[p] = -1e30f;
                continue;
            }
            px = px < 0.0 ? 0.0 : px;
            px = px > (double)(W - 1) ? (double)(W - 1) : px;
            py = py < 0.0 ? 0.0 : py;
            py = py > (double)(H - 1) ? (double)(H - 1) : py;
            safe_x[p] = (float)(int64_t)px;
            safe_y[p] = (float)(int64_t)py;
        }
}

/*
 * The same splat with ONE shared depth LUT (all channel LUTs of the literal loop are identical) and the winner map
 * exposed: winner[t] = source raster id, -1 hole, -2 hit-but-no-winner.  Sources whose target is out of range or
 * NaN (-1e30 marker) are skipped and counted in *dropped.  One frame.
 */
void oracle_splat_frame(const float* obj, const float* safe_y, const float* safe_x, const float* depth, int C, int H,
                        int W, float* out, float* valid, float* collision, int32_t* winner, float* dlut,
                        int64_t* dropped) {
    const size_t hw = (size_t)H * W;
    int64_t drop = 0;
    for (size_t k = 0; k < hw; ++k) dlut[k] = DLUT_INIT, winner[k] = -1;
    for (int j = 0; j < H; ++j)
        for (int i = 0; i < W; ++i) {
            const size_t p = (size_t)j * W + i;
            const float xf = safe_x[p], yf = safe_y[p];
            if (!(xf > -1.0f && xf < (float)W && yf > -1.0f && yf < (float)H)) {
                ++drop;
                continue;
            }
            const size_t t = (size_t)(int)yf * W + (int)xf;
            if (depth[p] < dlut[t]) {
                dlut[t] = depth[p];
                winner[t] = (int32_t)p;
            } else if (winner[t] == -1) {
                winner[t] = -2;
            }
        }
    for (size_t t = 0; t < hw; ++t) {
        const int32_t w = winner[t];
        valid[t] = (w != -1) ? 1.0f : 0.0f;
        collision[t] = (w == -2) ? 1.0f : 0.0f;
        for (int c = 0; c < C; ++c) out[(size_t)c * hw + t] = (w >= 0) ? obj[(size_t)c * hw + w] : 0.0f;
    }
    if (dropped) *dropped += drop;
}

/* utils.fix_warped_depth, utils.py:123-126 */
static inline float fix_depth(float d) {
    if (d == 0.0f) d = 100.0f;
    if (d > 99.5f) d = 100.0f;
    return d;
}

/*
 * One flow pair of the frame pipeline, preprocess.py:356-365 (inpaint excluded), float32 data path:
 *   disp = sBf / depth                       Convert.depth_to_disparity  preprocess.py:239-246
 *   flow = (disp * -1, 0 * -1)               Convert.disparity_to_flow   preprocess.py:249-254
 *   obj  = img0 | depth0 | flow * -1         preprocess.py:358
 *   FW(obj, flow, depth0)                    alt_cuda/fw.py:19-59 + fw_cuda_kernel.cu
 *   img1, depth1, back_flow *= valid; depth1 = fix_warped_depth(depth1)   preprocess.py:362-365
 * Frames are independent: `nthreads` pthreads take frames b = tid, tid + nthreads, ... (each owns its scratch).
 * img0[B,3,H,W], depth0[B,1,H,W], sBf[B] -> img1[B,3,H,W], depth1[B,1,H,W], back_flow[B,2,H,W], flow[B,2,H,W],
 * valid[B,1,H,W], collision[B,1,H,W].
 */
typedef struct {
    const float *img0, *depth0, *sBf;
    int B, H, W, tid, nthreads, rc;
    float *img1, *depth1, *back_flow, *flow, *valid, *collision;
} pair_job;

static void* pair_worker(void* arg) {
    pair_job* J = (pair_job*)arg;
    const int H = J->H, W = J->W;
    const size_t hw = (size_t)H * W;
    float* obj = (float*)malloc(sizeof(float) * 6 * hw);
    float* out = (float*)malloc(sizeof(float) * 6 * hw);
    float* sx = (float*)malloc(sizeof(float) * hw);
    float* sy = (float*)malloc(sizeof(float) * hw);
    float* dlut = (float*)malloc(sizeof(float) * hw);
    int32_t* win = (int32_t*)malloc(sizeof(int32_t) * hw);
    if (!obj || !out || !sx || !sy || !dlut || !win) {
        J->rc = -2;
    } else {
        for (int b = J->tid; b < J->B; b += J->nthreads) {
            const float* d = J->depth0 + (size_t)b * hw;
            float* f = J->flow + (size_t)b * 2 * hw;
            for (size_t p = 0; p < hw; ++p) {
                const float disp = J->sBf[b] / d[p];
                f[p] = disp * -1.0f;
                f[hw + p] = 0.0f * -1.0f;
            }
            memcpy(obj, J->img0 + (size_t)b * 3 * hw, sizeof(float) * 3 * hw);
            memcpy(obj + 3 * hw, d, sizeof(float) * hw);
            for (size_t p = 0; p < 2 * hw; ++p) obj[4 * hw + p] = f[p] * -1.0f;
            oracle_fw_targets_f32(f, H, W, sx, sy);
            float* v = J->valid + (size_t)b * hw;
            oracle_splat_frame(obj, sy, sx, d, 6, H, W, out, v, J->collision + (size_t)b * hw, win, dlut, NULL);
            for (int c = 0; c < 3; ++c)
                for (size_t p = 0; p < hw; ++p) J->img1[((size_t)b * 3 + c) * hw + p] = out[c * hw + p] * v[p];
            for (size_t p = 0; p < hw; ++p) J->depth1[(size_t)b * hw + p] = fix_depth(out[3 * hw + p] * v[p]);
            for (int c = 0; c < 2; ++c)
                for (size_t p = 0; p < hw; ++p)
                    J->back_flow[((size_t)b * 2 + c) * hw + p] = out[(4 + c) * hw + p] * v[p];
        }
    }
    free(obj), free(out), free(sx), free(sy), free(dlut), free(win);
    return NULL;
}

int oracle_disparity_pair(const float* img0, const float* depth0, const float* sBf, int B, int H, int W, float* img1,
                          float* depth1, float* back_flow, float* flow, float* valid, float* collision, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    pair_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        pair_job j = {img0, depth0, sBf, B, H, W, t, nthreads, 0, img1, depth1, back_flow, flow, valid, collision};
        jobs[t] = j;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, pair_worker, &jobs[t]);
    pair_worker(&jobs[0]);
    int rc = jobs[0].rc;
    for (int t = 1; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    return rc;
}

/* Batched FW.forward (fw.py prologue + splat), frames in sequence. */
int oracle_fw_batch_f32(const float* obj, const float* flow, const float* depth, int B, int C, int H, int W, float* out,
                        float* valid, float* collision, int32_t* winner, int64_t* dropped) {
    const size_t hw = (size_t)H * W;
    int rc = 0;
    int64_t drop_total = 0;
    for (int b = 0; b < B; ++b) {
        float* sx = (float*)malloc(sizeof(float) * hw);
        float* sy = (float*)malloc(sizeof(float) * hw);
        float* dlut = (float*)malloc(sizeof(float) * hw);
        if (!sx || !sy || !dlut) {
            rc = -2;
        } else {
            int64_t drop = 0;
            oracle_fw_targets_f32(flow + (size_t)b * 2 * hw, H, W, sx, sy);
            oracle_splat_frame(obj + (size_t)b * C * hw, sy, sx, depth + (size_t)b * hw, C, H, W,
                               out + (size_t)b * C * hw, valid + (size_t)b * hw, collision + (size_t)b * hw,
                               winner + (size_t)b * hw, dlut, &drop);
            drop_total += drop;
        }
        free(sx), free(sy), free(dlut);
    }
    if (dropped) *dropped = drop_total;
    return rc;
}
