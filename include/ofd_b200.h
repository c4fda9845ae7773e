/*
 * ofd_b200.h — C ABI of libofd_b200.so, the B200-native (sm_100a) flow-synthesis hot path.
 *
 * This library replaces the reference's torch extension `fw_cuda`
 * (alt_cuda/fw_cuda.cpp:15-30, alt_cuda/fw_cuda_kernel.cu:10-83) and fuses the per-pixel work the
 * reference does in torch around it (alt_cuda/fw.py:19-59, preprocess.py:237-326, geometry.py:17-67,
 * utils.py:102-126, bilateral_filter.py:13-60).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - tensors are dense NCHW, batch-major, no strides; `B` frames of `H x W`, H*W < 2^31;
 *   - `stream` is a cudaStream_t (opaque here so the header needs no CUDA include);
 *   - every entry point returns 0 on success, a negative OFD_E_* code for a rejected call (nothing was
 *     launched) or a positive cudaError_t; ofd_last_error_string() explains the last failure on the
 *     calling thread.  Nothing throws across this ABI, launches are asynchronous on `stream`;
 *   - the library keeps no global mutable state; scratch ("key workspace") is caller-owned.
 *
 * Key workspace.  The z-buffer is a plane of packed 64-bit keys  (ordered_depth_bits << 32 | source_raster_id),
 * 8 bytes per target pixel.  It must hold the "untouched" pattern (all bytes 0xFF) when a splat starts; every
 * splat entry point leaves it in that state again when it returns success (the gather pass re-arms the keys it
 * consumed), so ofd_workspace_reset() is only needed once after allocation or after a failed call.
 */
#ifndef OFD_B200_H_
#define OFD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ofd_stream_t; /* cudaStream_t */

/* ---- return codes ------------------------------------------------------------------------------------- */
#define OFD_OK 0
#define OFD_E_NULL (-1)      /* a required pointer is NULL                       */
#define OFD_E_SHAPE (-2)     /* B/C/H/W out of the supported range               */
#define OFD_E_DTYPE (-3)     /* unsupported dtype code                           */
#define OFD_E_ARG (-4)       /* any other invalid argument (epilogue, window...) */
#define OFD_E_WORKSPACE (-5) /* workspace too small / misaligned                 */

/* dtype codes */
#define OFD_F32 0
#define OFD_F64 1

/* gather epilogues (ofd_splat_flow) */
#define OFD_EPI_NONE 0   /* alt_cuda/fw.py:19-59: raw FW.forward result                                   */
#define OFD_EPI_CONCAT 1 /* preprocess.py:307-313 ConcatFlow: out = (warp + aux) * valid                  */
#define OFD_EPI_BACK 2   /* preprocess.py:321-326 BackFlow:   out = (warp * -1) * valid                   */

#define OFD_MAX_CHANNELS 8

/* counters written by the gather pass when a counter block is supplied (uint64 each, accumulated) */
#define OFD_CNT_HIT 0       /* target pixels with valid == 1                                  */
#define OFD_CNT_HOLE 1      /* target pixels with valid == 0                                  */
#define OFD_CNT_COLLISION 2 /* target pixels with collision == 1                              */
#define OFD_CNT_DROPPED 3   /* sources dropped: NaN flow / out-of-range explicit target       */
#define OFD_CNT_TIE_SRC 4   /* sources that tie the winning depth but lose on raster index (deterministic in the
                               reference's serial loop and here; counted for information by a census pass) */
#define OFD_CNT_FRAMES 5
#define OFD_CNT_PAIRS 6
#define OFD_CNT_SLOTS 8

int ofd_version(void);
const char* ofd_last_error_string(void);

/* Bytes of key workspace a splat over B frames of HxW needs. */
size_t ofd_workspace_bytes(int B, int H, int W);
/* Fill a fresh (or possibly dirty) workspace with the "untouched" key pattern. */
int ofd_workspace_reset(void* ws, size_t bytes, ofd_stream_t stream);

/*
 * ofd_splat_targets — the exact contract of fw_cuda.forward_warping (alt_cuda/fw_cuda.cpp:15-26):
 *   obj[B,C,H,W], safe_y/safe_x/depth[B,1,H,W] of one dtype -> out[B,C,H,W], valid/collision[B,1,H,W].
 *   Target of source (j,i) is (int)safe_y, (int)safe_x (float->int truncation of the accessor index,
 *   fw_cuda_kernel.cu:31-32).  Winner per target = min depth among sources with depth < 1000, ties -> lowest
 *   raster index (the serial loop order, :28-29,34); valid = any hit; collision = hit but no winner (:39-44).
 *   Divergence (documented): a target outside [0,H)x[0,W) or NaN is undefined behaviour in the reference;
 *   here that source is dropped and counted in OFD_CNT_DROPPED.
 *   `winner` (optional, int32 [B,1,H,W]) receives the winning source raster id, -1 for holes, -2 for collisions.
 *   dtype: OFD_F32, or OFD_F64 (all four tensors and the three outputs double, as AT_DISPATCH_FLOATING_TYPES at
 *   fw_cuda_kernel.cu:70 allows; two key planes, so the workspace must hold 2 * ofd_workspace_bytes(B,H,W)).
 */
int ofd_splat_targets(const void* obj, const void* safe_y, const void* safe_x, const void* depth, int dtype,
                      int B, int C, int H, int W, void* out, void* valid, void* collision, int32_t* winner,
                      uint64_t* counters, void* ws, size_t ws_bytes, ofd_stream_t stream);

/*
 * ofd_splat_flow — FW.forward (alt_cuda/fw.py:19-59) with the torch prologue fused into the z-test:
 *   target = trunc(clamp((i,j) + flow, 0, (W-1,H-1))) evaluated in the flow's dtype (fw.py:31,37-42);
 *   flow is [B,2,H,W] float32 (flow_dtype OFD_F32) or float64 (OFD_F64); obj/depth/out are float32.
 *   Divergence (documented): NaN flow is UB in the reference; here the source is dropped and counted.
 *   epilogue: OFD_EPI_*; `aux` is flowAB[B,2,H,W] for OFD_EPI_CONCAT (C must be 2), else NULL.
 */
int ofd_splat_flow(const float* obj, const void* flow, int flow_dtype, const float* depth, int B, int C, int H,
                   int W, float* out, float* valid, float* collision, int32_t* winner, int epilogue,
                   const float* aux, uint64_t* counters, void* ws, size_t ws_bytes, ofd_stream_t stream);
/*
 * ofd_splat_flow_rows — ofd_splat_flow for a HORIZONTAL warp flow: the caller guarantees flow[:,1] == +-0 everywhere (it is not read), as
 * the pipeline's disparity flows flow01 / back_flow01 are by construction (preprocess.py:253,361-363) - the warp flow of the two ConcatFlow
 * splats of a frame group (preprocess.py:400,414).  Sources then stay in their row: the z-buffer of a row lives in shared memory (two
 * 32-bit shared-memory atomicMin passes = the serial loop's winner), no global atomics, no key workspace, ONE launch, 40 B/px instead of
 * 72 for a ConcatFlow.  float32 only, C == 2 (flow payloads), W <= 2048; same results as ofd_splat_flow on such flows (tested bit for bit).
 * valid_mul (nullable, [B,1,H,W]): the valid plane is written as valid * valid_mul (preprocess.py:415: flow13_valid * img1_valid); the
 * epilogue's own use of valid (ConcatFlow / BackFlow masking, :312,325) is unaffected.
 */
int ofd_splat_flow_rows(const float* obj, const float* flow, const float* depth, int B, int C, int H, int W, float* out,
                        float* valid, float* collision /*nullable*/, int epilogue, const float* aux, const float* valid_mul /*nullable*/,
                        ofd_stream_t stream);

/*
 * ofd_disparity_flow — Convert.depth_to_disparity + disparity_to_flow(random_sign=False)
 * (preprocess.py:239-254): flow = (-(sBf / depth), -0.0), sBf = float32(s*B*f) promoted to the depth dtype.
 *   depth[B,1,H,W] (depth_dtype f32|f64) -> flow[B,2,H,W] of the same dtype; sBf[B] float32 on the DEVICE.
 */
int ofd_disparity_flow(const void* depth, int depth_dtype, const float* sBf, int B, int H, int W, void* flow,
                       ofd_stream_t stream);

/*
 * ofd_disparity_pair — one whole "flow pair" of preprocess.py:355-366 (minus inpaint) in ONE kernel:
 *   flow01 = (-(sBf/depth0), -0.0); splat of obj = img0(3) | depth0(1) | -flow01(2) along flow01 z-tested
 *   with depth0; results masked by valid; fix_warped_depth (utils.py:123-126) on the warped depth.
 *   Because flow.y == -0.0 exactly, every source stays in its row: the z-buffer lives in shared memory,
 *   no global atomics and no key workspace.
 *   in : img0[B,3,H,W] f32, depth0[B,1,H,W] (depth_dtype f32|f64), sBf[B] (device, float32)
 *   out: img1[B,3,H,W], depth1[B,1,H,W], back_flow[B,2,H,W], flow[B,2,H,W] (f32; NULL to skip),
 *        valid[B,1,H,W], collision[B,1,H,W] (NULL to skip).
 */
int ofd_disparity_pair(const float* img0, const void* depth0, int depth_dtype, const float* sBf, int B, int H,
                       int W, float* img1, float* depth1, float* back_flow, float* flow, float* valid,
                       float* collision, uint64_t* counters, ofd_stream_t stream);

/*
 * ofd_disparity_pair_ragged — ofd_disparity_pair over a RAGGED batch (BASELINE config 2: mixed-resolution frames, every frame
 * synthesised exactly as the single-frame call would): image i is H_host[i] x W_host[i] and starts at PIXEL offset
 * offset_host[i] of the packed buffers - a C-channel tensor stores it densely as [C,H_i,W_i] at element C * offset_host[i]
 * (depth0 / depth1 / valid / collision: C = 1, img0 / img1: 3, back_flow / flow: 2); sBf[n_images] on the device.
 * One persistent launch per 96 frames for any H, W <= ~2890 and offsets (work units may start on any pixel, every plane on any
 * 16-byte phase; the buffers themselves must be 16-byte aligned); only the frame that ends the buffers on an odd 16-byte
 * boundary, and rows wider than that, take one launch per frame.  H_host / W_host / offset_host are HOST arrays.
 */
int ofd_disparity_pair_ragged(const float* img0, const void* depth0, int depth_dtype, const float* sBf, int n_images,
                              const int* H_host, const int* W_host, const size_t* offset_host, float* img1, float* depth1,
                              float* back_flow, float* flow /*nullable*/, float* valid, float* collision /*nullable*/,
                              uint64_t* counters, ofd_stream_t stream);

/*
 * ofd_reproject_flow — Convert.depth_to_random_flow (preprocess.py:265-298) = geometry.BackprojectDepth.forward
 * (geometry.py:37-42) + geometry.Project3D.forward (geometry.py:56-67) + de-normalisation, per pixel, fused:
 *   ray = invK3 (x,y,1);  X = depth*ray (float32);  c = P (X,1);  u = c.x/(c.z+eps), v = c.y/(c.z+eps);
 *   u <- ((u/(W-1) - .5)*2 + 1)/2*(W-1)  (same for v with H);  flow = (u - x, v - y).
 *   cam: per frame 21 floats on the DEVICE = invK3 row-major (9) followed by P = (K T)[:3,:] row-major (12).
 *   depth[B,1,H,W] f32|f64 (depth*ray is evaluated in the depth dtype then rounded to f32, geometry.py:39-40).
 */
int ofd_reproject_flow(const void* depth, int depth_dtype, const float* cam, float eps, int B, int H, int W,
                       float* flow, ofd_stream_t stream);

/*
 * Class-level geometry entry points, kept so geometry.py's modules stay drop-in (the fused ofd_reproject_flow is
 * the fast path).  ofd_backproject = BackprojectDepth.forward (geometry.py:37-42): depth[B,1,H,W] (f32|f64),
 * invk3[B,9] -> cam_points[B,4,H*W] f32.  ofd_project = Project3D.forward (geometry.py:56-67): cam_points, P[B,12]
 * -> pix_coords[B,H,W,2] in [-1,1] and z[B,1,H*W].  All pointers are device pointers.
 */
int ofd_backproject(const void* depth, int depth_dtype, const float* invk3, int B, int H, int W, float* cam_points,
                    ofd_stream_t stream);
int ofd_project(const float* cam_points, const float* P, float eps, int B, int H, int W, float* pix_coords, float* z,
                ofd_stream_t stream);

/*
 * ofd_frame_splat — one image+flow splat of the frame pipeline (preprocess.py:372-382 / 385-394 / 401-411):
 *   obj = img(3) | depth(1) | -flow(2) | valid_in(1, optional) gathered from separate planes (no torch.cat);
 *   z-test with `depth` along `flow` (float32 [B,2,H,W]); epilogue: valid' = valid * warp(valid_in) (or valid),
 *   all outputs * valid', fix_warped_depth on the warped depth.
 *   out: img_out[B,3,H,W], depth_out[B,1,H,W], back_flow[B,2,H,W], valid_out[B,1,H,W] (= valid'),
 *        collision[B,1,H,W] (NULL to skip), raw_valid[B,1,H,W] (NULL to skip).
 */
int ofd_frame_splat(const float* img, const float* depth, const float* flow, const float* valid_in, int B, int H,
                    int W, float* img_out, float* depth_out, float* back_flow, float* valid_out, float* collision,
                    float* raw_valid, uint64_t* counters, void* ws, size_t ws_bytes, ofd_stream_t stream);

/*
 * ofd_concat_frame_splat — ConcatFlow along a HORIZONTAL warp flow fused with the frame splat that consumes its result: the pairs 0->2'
 * and 1->3' of a frame group (preprocess.py:400-411 and :414-424).
 *   flowAC       = (FW(flowBC, warp_flow, depthB) + flowAB) * valid        ConcatFlow.forward, :307-313  (warp_flow.y == +-0, not read)
 *   flowAC_valid = valid [* valid_mul]                                       (:415)
 *   img_out, depth_out, back_flow, valid_out[, collision] = the frame splat of (img, depth_src, -flowAC, flowAC_valid) along flowAC by
 *   depth_src, as ofd_frame_splat with valid_in = flowAC_valid                (:401-411)
 * Two launches instead of three: the row-local ConcatFlow kernel also reduces the packed keys of the frame splat's z-test (sources = the
 * pixels of the row it just produced), then the gather runs.  Same results as ofd_splat_flow_rows + ofd_frame_splat, bit for bit.
 * float32, W % 4 == 0, W <= 2048, planes 16-byte aligned; ws as for ofd_frame_splat.
 */
int ofd_concat_frame_splat(const float* flowBC, const float* warp_flow, const float* depthB, const float* flowAB,
                           const float* valid_mul /*nullable*/, const float* img, const float* depth_src, int B, int H, int W,
                           float* flowAC, float* flowAC_valid, float* img_out, float* depth_out, float* back_flow, float* valid_out,
                           float* collision /*nullable*/, uint64_t* counters /*nullable*/, void* ws, size_t ws_bytes,
                           ofd_stream_t stream);
/* The same splat when the warp flow is float64 (the dataset path: a flow composed with the float64 disparity flow,
 * preprocess.py:400-401 with cv2-loaded depth): targets are evaluated in float64 from `flow` (alt_cuda/fw.py:31,37-42), the
 * payload channels 4-5 are -flow_payload, the caller's float32 rounding of the same flow (fw.py:45 casts obj to float32). */
int ofd_frame_splat_f64(const float* img, const float* depth, const double* flow, const float* flow_payload,
                        const float* valid_in, int B, int H, int W, float* img_out, float* depth_out, float* back_flow,
                        float* valid_out, float* collision, float* raw_valid, uint64_t* counters, void* ws,
                        size_t ws_bytes, ofd_stream_t stream);

/*
 * ofd_reproject_pair — a whole 6-DoF "flow pair" (preprocess.py:372-382 / 385-394 minus inpaint) in two launches:
 * the z-test computes the reprojection flow from `depth` and the camera constants in place and writes it out once as
 * `flow_out`; the gather pulls image, depth, -flow and valid_in of the winning source.  Same results as
 * ofd_reproject_flow followed by ofd_frame_splat, 8 B/px less HBM traffic and one launch fewer.
 * cam[B,21] on the device as for ofd_reproject_flow.
 */
int ofd_reproject_pair(const float* img, const float* depth, const float* cam, float eps, const float* valid_in, int B,
                       int H, int W, float* img_out, float* depth_out, float* back_flow, float* flow_out,
                       float* valid_out, float* collision, float* raw_valid, uint64_t* counters, void* ws,
                       size_t ws_bytes, ofd_stream_t stream);

/*
 * ofd_normalize_depth — utils.normalize_depth (utils.py:102-116) per frame: 0 -> 100, >100 -> 100, min over
 * all, 100 -> 0, max, affine map to [1,99], formerly-invalid pixels -> 100.  Out of place (the reference's
 * half-mutation of its argument is not reproduced).  scratch: 2*B uint64 words (device).
 */
int ofd_normalize_depth(const void* depth, int dtype, int B, int H, int W, void* out, void* scratch,
                        ofd_stream_t stream);

/*
 * ofd_normalize_depth_ragged — ofd_normalize_depth for a RAGGED batch (BASELINE config 2): image i is count_host[i] elements
 * at element offset offset_host[i] of depth / out (HOST arrays); each image gets its own min / max.  3 launches per 128
 * images.  scratch: 2*n_images uint64 words (device).
 */
int ofd_normalize_depth_ragged(const void* depth, int dtype, int n_images, const size_t* count_host,
                               const size_t* offset_host, void* out, void* scratch, ofd_stream_t stream);

/*
 * ofd_depth_from_png — the arithmetic of the reference's depth loaders after cv2.imread, on the device, so a
 * depth map crosses PCIe as its 8/16-bit PNG payload instead of float64:
 *   OFD_SRC_RELDEPTH  = utils.get_depth(smooth=True) (utils.py:47-59) with utils.smooth_closer (utils.py:118-121):
 *                       depth = 1 / (255 - min(v, 240))
 *   OFD_SRC_DISPARITY = utils.get_disparity (utils.py:61-72) + Convert.disparity_to_depth (preprocess.py:257-262):
 *                       depth = (1 / (v * 63 / 255 + 0.005)) * 50   (torch evaluates int / tensor as reciprocal * int)
 * evaluated in float64 in numpy's / torch's op order; depth_dtype OFD_F64 (what the reference holds) or OFD_F32 (that value
 * rounded once).  src: n samples of src_bits (8 or 16) bits; the result feeds ofd_normalize_depth.
 */
#define OFD_SRC_RELDEPTH 0
#define OFD_SRC_DISPARITY 1
int ofd_depth_from_png(const void* src, int src_bits, int kind, size_t n, void* depth, int depth_dtype,
                       ofd_stream_t stream);

/* ofd_fix_warped_depth — utils.fix_warped_depth (utils.py:123-126), in place: 0 -> 100, > 99.5 -> 100. */
int ofd_fix_warped_depth(float* depth, size_t n, ofd_stream_t stream);

/*
 * ofd_inpaint_mask — the hole-mask logic of utils.inpaint (utils.py:137-149; SURVEY 8f-1, the step after every image
 * splat): M = (valid != collision); M' = 3x3 dilate(M); H' = valid * (M' == M); mask = 1 - H' as uint8 [B,1,H,W],
 * i.e. the mask the reference hands to cv2.inpaint (the Telea fill itself stays a host-side hook).
 */
int ofd_inpaint_mask(const float* valid, const float* collision, int B, int H, int W, uint8_t* mask,
                     ofd_stream_t stream);

/*
 * ofd_special_flow — SpecialFlow.forward (preprocess.py:24-105): analytic augmentation flows.
 *   kind 5 flip (vertical, :47-60; no params), 6 rotate (:62-79), 7 shear (:81-99).
 *   params_host (HOST, 10 floats, kinds 6/7): cx, cy, M row-major (4), Mrev row-major (4) with
 *   p1 = (p0 - c) @ M + c, p_prev = (p0 - c) @ Mrev + c; kind 7 ignores the centre (p1 = p0 @ M).
 *   out: flow[2,H,W] = p1 - p0 and back_flow[2,H,W] = p_prev - p0 (float32).
 */
int ofd_special_flow(int kind, const float* params_host, int H, int W, float* flow, float* back_flow,
                     ofd_stream_t stream);

/*
 * ofd_special_flow_batch — one special flow per sample of a batch in one launch (per 48 samples): kinds_host[B] (HOST
 * ints, 5/6/7) and params_host[B,10] (HOST floats, layout as ofd_special_flow; ignored for kind 5)
 * -> flow[B,2,H,W], back_flow[B,2,H,W].
 */
int ofd_special_flow_batch(const int* kinds_host, const float* params_host, int B, int H, int W, float* flow,
                           float* back_flow, ofd_stream_t stream);

/*
 * ofd_augment_pairs — the geometric branch of augment_flow (preprocess.py:116-147, inpaint excluded) for a batch of B
 * pairs (img0, depth0, img1, depth1, flow01, back_flow01), one special flow per sample (BASELINE config 4: in-loop
 * augmentation for RAFT training).  13 launches issued back to back from this one call:
 *   special/back_special   = SpecialFlow(kind_b, params_b)                                   (:118)
 *   aug0_flow              = ConcatFlow(back_special, special, flow01, depth0)               (:121)
 *   aug1_flow              = ConcatFlow(flow01, back_flow01, special, depth1)                (:122)
 *   aug_img0 | aug_depth0  = FW(img0 | depth0, special, depth0), fix_warped_depth on depth   (:124-129)
 *   aug_img1 | aug_depth1  = FW(img1 | depth1, special, depth1), fix_warped_depth on depth   (:130-135)
 *   back_aug0_flow         = BackFlow(aug0_flow, aug_depth0)                                 (:137)
 *   back_aug1_flow         = BackFlow(aug1_flow, depth0)                                     (:138)
 * valid_img* / collision_img* [B,1,H,W] are the masks of the two image warps (what utils.inpaint is handed, :127,133;
 * collision nullable); scratch_valid [B,1,H,W] receives the (unused) masks of the four flow splats.
 */
int ofd_augment_pairs(const float* img0, const float* depth0, const float* img1, const float* depth1,
                      const float* flow01, const float* back_flow01, const int* kinds_host, const float* params_host,
                      int B, int H, int W, float* special_flow, float* back_special_flow, float* aug_img0,
                      float* aug_depth0, float* aug0_flow, float* back_aug0_flow, float* aug_img1, float* aug_depth1,
                      float* aug1_flow, float* back_aug1_flow, float* valid_img0, float* collision_img0,
                      float* valid_img1, float* collision_img1, float* scratch_valid, uint64_t* counters, void* ws,
                      size_t ws_bytes, ofd_stream_t stream);

/*
 * ofd_bilateral_iter — one iteration of sparse_bilateral_filtering (bilateral_filter.py:33-58):
 *   discontinuity map from `depth_in` (|1/d - 1/d'| > thr on 4-neighbours of the interior, :63-116), forced to 1
 *   where depth_orig == 0 (:46), border ring edge-replicated (:141-147), then the gated median of window x window
 *   (:167-198) with the reference's float32 cumsum rank rule.  dtype f32|f64 applies to all three planes.
 */
int ofd_bilateral_iter(const void* depth_in, const void* depth_orig, int dtype, int H, int W, int window,
                       double threshold, void* depth_out, ofd_stream_t stream);

/*
 * ofd_bilateral_iter_masked — one iteration with the reference's BINARY mask (bilateral_filter.py:48-49,72-80,161,169-170,180-182):
 * a neighbour difference counts only between two unmasked pixels; masked pixels are never discontinuities and keep their depth;
 * masked taps and taps outside the image (the mask is zero-padded, not ring-replicated) are left out of the median.
 * mask: uint8 [H,W] on the device, 0 = masked.  mask_coef_f64: 1 when the reference's mask array is float64 or integer (its median
 * coefficients, float32 * mask.dtype, are then float64 and the rank rule k(n) runs in float64), 0 for uint8 / bool / float32 masks.
 * Windows 3, 5, 7.
 */
int ofd_bilateral_iter_masked(const void* depth_in, const void* depth_orig, const uint8_t* mask, int mask_coef_f64, int dtype,
                              int H, int W, int window, double threshold, void* depth_out, ofd_stream_t stream);

/*
 * ofd_bilateral_iter_batch — the same iteration over a RAGGED batch (BASELINE config 2: mixed-resolution frames, each
 * filtered independently exactly as ofd_bilateral_iter would): image i is H_host[i] x W_host[i], stored densely at element
 * offset offset_host[i] of depth_in / depth_orig / depth_out (the three buffers share one layout).  One launch per 64
 * images; H_host / W_host / offset_host are HOST arrays.
 */
int ofd_bilateral_iter_batch(const void* depth_in, const void* depth_orig, int dtype, int n_images, const int* H_host,
                             const int* W_host, const size_t* offset_host, int window, double threshold,
                             void* depth_out, ofd_stream_t stream);

/*
 * Host-buffer front end of the flow-pair path (what a reference-side caller holding numpy arrays or pinned CPU
 * tensors binds; the reference crosses host<->device around every FW call, preprocess.py:350-366,437-447).
 * A pipeline owns three device staging slots with one stream each: chunk k does host->device copies, the pair
 * kernel and device->host copies on slot k%3, so both copy engines and the SMs overlap.  Host buffers should be
 * page-locked for the copies to overlap.  `run` returns after every result byte has landed in the host buffers.  The y planes of
 * flow and back_flow are constants of the virtual-stereo pair (-0.0 / +0.0): they are not transferred (8 of the 40 result bytes per
 * pixel) but written into the host buffers by host threads while the copies run; valid / collision (0.0f / 1.0f planes) cross PCIe as
 * one packed byte per pixel and are expanded into the caller's float planes by the same threads (25 instead of 40 B/px on the wire,
 * H*W a multiple of 4) - the buffers end up complete and bit-identical either way.  img1 is a selection of img0's pixels: when a byte
 * carries every value of a chunk exactly (uint8-valued frames - what the reference's loader delivers, utils.py:17-25), its three planes
 * cross as bytes too (16 instead of 25 B/px on the wire) and are widened by the host threads; the packing kernel verifies every chunk on
 * the device and a chunk that fails is redone with float planes inside the same call (the pipeline then stops trying).  On by default
 * unless more than two ranks share the node (LOCAL_WORLD_SIZE > 2: host memory, not PCIe, is the limit there); OFD_HOST_IMG_BYTES=0/1.
 * Host threads: OFD_HOST_WORKERS (default: cores of the creating thread's affinity mask minus one, clamped to 2..4; read when the pipeline is
 * CREATED) persistent threads per pipeline; they sleep on a
 * condition variable between runs and chunks and inherit the CPU affinity of the creating thread.  OFD_HOST_MASK_BYTES=0 (read at
 * creation) sends the masks as float planes.
 * ofd_pair_pipeline_run_flags: the same call with option bits.  OFD_PIPE_KEEP_CONST_PLANES: the caller recycles result buffers
 * whose flow.y / back_flow.y planes already hold -0.0 / +0.0 (e.g. from an earlier run into the same buffers); the pipeline then
 * never touches those planes (8 B/px fewer host stores per pair).
 */
#define OFD_PIPE_KEEP_CONST_PLANES 1u
typedef struct ofd_pair_pipeline ofd_pair_pipeline;
int ofd_pair_pipeline_create(int device, int H, int W, int chunk_frames, ofd_pair_pipeline** out);
int ofd_pair_pipeline_run(ofd_pair_pipeline* p, const float* img0_host, const float* depth0_host,
                          const float* sBf_host, int B, float* img1_host, float* depth1_host, float* back_flow_host,
                          float* flow_host /*nullable*/, float* valid_host, float* collision_host /*nullable*/);
int ofd_pair_pipeline_run_flags(ofd_pair_pipeline* p, const float* img0_host, const float* depth0_host,
                                const float* sBf_host, int B, float* img1_host, float* depth1_host, float* back_flow_host,
                                float* flow_host /*nullable*/, float* valid_host, float* collision_host /*nullable*/,
                                unsigned flags);
/*
 * Input side on the device.
 * ofd_jpeg_*: utils.get_img (utils.py:17-25) decodes on the host with cv2.imread(path, -1) and converts BGR uint8 HWC to float32 CHW;
 * here the compressed JPEG bytes (ReDWeb's images, dataloader.py:28) go to nvJPEG (loaded with dlopen at the first create; the library has
 * no link-time dependency on it), which decodes on the GPU, and a kernel widens the planar B | G | R bytes to out_bgr[3,H,W] float32 - a
 * frame crosses PCIe once, as compressed bytes.  JPEG decoders are not bit-specified: the result is within a few grey levels of cv2's
 * (libjpeg-turbo), see tests.  ofd_jpeg_info reads H, W (and the component count) from the header; ofd_jpeg_decode checks them against the
 * tensor it is given.  data_host is HOST memory.  PNG (deflate) stays on the host.
 * ofd_resize_bilinear_aa: T.Resize(size)(depth) of dataloader.py:31-32,57-58 = torchvision's antialiased bilinear resize (ATen
 * upsample_bilinear2d_aa, align_corners = false) of src[B,H,W] (f32|f64) to dst[B,H_out,W_out]; tmp holds B*H*W_out elements when both axes
 * change.  Bit-identical to torchvision when upscaling, within 2 ulp when downscaling.
 */
typedef struct ofd_jpeg_decoder ofd_jpeg_decoder;
int ofd_jpeg_decoder_create(int device, ofd_jpeg_decoder** out);
int ofd_jpeg_info(ofd_jpeg_decoder* d, const uint8_t* data_host, size_t nbytes, int* H, int* W, int* components);
int ofd_jpeg_decode(ofd_jpeg_decoder* d, const uint8_t* data_host, size_t nbytes, float* out_bgr, int H, int W,
                    ofd_stream_t stream);
void ofd_jpeg_decoder_destroy(ofd_jpeg_decoder* d);
int ofd_resize_bilinear_aa(const void* src, int dtype, int B, int H, int W, int H_out, int W_out, void* dst, void* tmp,
                           ofd_stream_t stream);

/*
 * ofd_inpaint_telea — the fill of utils.inpaint (utils.py:136-151: cv2.inpaint(img_u8, mask, 3, cv2.INPAINT_TELEA) on the host, 95
 * calls per frame in preprocess.py) on the device, for a batch: img[B,3,H,W] float32 (truncated to uint8 like .astype(np.uint8),
 * utils.py:147), mask[B,1,H,W] uint8 (!= 0 = fill; ofd_inpaint_mask produces it), range = inpaint radius (the reference uses 3)
 * -> out[B,3,H,W] float32 (uint8-valued).  Telea's fast-marching method with OpenCV's per-pixel arithmetic, marched in LAYERS
 * (all hole pixels with a 4-neighbour in earlier layers at once) instead of OpenCV's serial heap order: not bit-identical to
 * cv2.inpaint by construction - tests state and measure the difference (fraction of bytes off by more than one grey level).
 * One cooperative kernel; `ws` holds ofd_inpaint_workspace_bytes(B,H,W) bytes (256-byte aligned, no initialisation needed).
 * stats_host (optional, HOST pointer to 2 x uint32): layers marched and pixels filled; passing it synchronises the stream.
 */
size_t ofd_inpaint_workspace_bytes(int B, int H, int W);
int ofd_inpaint_telea(const float* img, const uint8_t* mask, int B, int H, int W, int range, float* out, void* ws,
                      size_t ws_bytes, uint32_t* stats_host, ofd_stream_t stream);

/*
 * ofd_plane_ops — a table of plane operations dst[0..hw) = f(src) in one launch: how a batch of training samples is put together
 * from the planes of a frame group on the device (BASELINE config 4).  Replaces, per sample, the reader's channel slicing of the
 * pre-baked arrays (dataloader.py:93-126: img0 / depth / img1 / flow of pair group 0..2) and the writer's photometric functions
 * (preprocess.py:150-163): OFD_PLANE_SCALE = img * scale (type 0), OFD_PLANE_ADD = img[channel] + shift (type 1), OFD_PLANE_GRAY =
 * (r * 0.2989 + g * 0.5870) + b * 0.1140 with r, g, b = src, src + gray_stride, src + 2 * gray_stride (type 2), OFD_PLANE_COPY.
 * float32 operations in exactly that order (no FMA contraction): the values torch's elementwise kernels give.
 *   ops_dev: DEVICE table of n_ops entries (8-byte aligned); aligned16 != 0 promises that every src / dst is 16-byte aligned
 *   (with hw % 4 == 0 the planes then move as float4).  A plane must not be both a source and a destination of the same table.
 */
#define OFD_PLANE_COPY 0
#define OFD_PLANE_SCALE 1
#define OFD_PLANE_ADD 2
#define OFD_PLANE_GRAY 3
typedef struct ofd_plane_op {
    const float* src; /* device plane of hw floats (GRAY: the first of three planes gray_stride floats apart) */
    float* dst;       /* device plane of hw floats */
    int op;           /* OFD_PLANE_* */
    float p;          /* scale / shift */
} ofd_plane_op;
int ofd_plane_ops(const ofd_plane_op* ops_dev, int n_ops, size_t hw, size_t gray_stride, int aligned16, ofd_stream_t stream);

/*
 * ofd_copy_rows_to_host — device -> host copy of `rows` rows of `width_bytes` with independent pitches (one asynchronous
 * cudaMemcpy2DAsync on `stream`).  The sweep uses it to scatter each result tensor [B,c,H,W] of a frame group straight into
 * its channel slice of the page-locked [B,44,H,W] group array (preprocess.py:437-447 layout): rows = B, width = c*H*W*4,
 * dst pitch = 44*H*W*4 - the 44-channel concatenation the reference builds with torch.cat never exists on the device.
 */
int ofd_copy_rows_to_host(const void* src, size_t src_pitch_bytes, void* dst_host, size_t dst_pitch_bytes,
                          size_t width_bytes, size_t rows, ofd_stream_t stream);
/* Lossless byte transport of uint8-valued float planes (the image channels of a frame group): ofd_pack_u8 narrows n device floats to bytes
 * and raises *flag (device int, zeroed by the caller) if any value is not exactly a uint8 (sign of zero included); ofd_host_widen_u8 widens n
 * HOST bytes into HOST floats with non-temporal stores on the calling thread.  sweep.PinnedGroupSink uses the pair: 18 of the 44 channels
 * of a group cross PCIe at 1 B instead of 4 and a failed check falls back to the float planes. */
int ofd_pack_u8(const float* src, uint8_t* dst, size_t n, int* flag, ofd_stream_t stream);
int ofd_host_widen_u8(const uint8_t* src_host, size_t n, float* dst_host);
/* Host-memory probe used by tools/probe_pcie.py: streams n floats of `value` into dst_host with non-temporal stores on the
 * calling thread (the store pattern of the pipeline's host threads); no CUDA call. */
int ofd_host_stream_fill(float* dst_host, size_t n, float value);
/* Compact transport of the same pipeline: colour planes and the valid / collision masks cross PCIe as uint8 and the
 * two constant planes (flow.y == -0.0, back_flow.y == +0.0) are not transferred: 7 B/px up, 16-17 B/px down instead
 * of 16 and 40.  Lossless when img0 holds integers 0..255 - what the reference's loader delivers (cv2.imread then
 * .type(float32), utils.py:17-25).  back_flow_x / flow_x are [B,1,H,W]. */
int ofd_pair_pipeline_run_u8(ofd_pair_pipeline* p, const uint8_t* img0_u8_host, const float* depth0_host,
                             const float* sBf_host, int B, uint8_t* img1_u8_host, float* depth1_host,
                             float* back_flow_x_host, float* flow_x_host /*nullable*/, uint8_t* valid_u8_host,
                             uint8_t* collision_u8_host /*nullable*/);
void ofd_pair_pipeline_destroy(ofd_pair_pipeline* p);

#ifdef __cplusplus
}
#endif
#endif /* OFD_B200_H_ */
