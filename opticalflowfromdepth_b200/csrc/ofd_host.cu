// ofd_host.cu — host-buffer front end of the flow-pair path: what a caller holding numpy / pinned CPU tensors
// binds (the reference moves every frame host->device->host around its FW call, preprocess.py:350-366,437-447).
// A pipeline owns NSLOT device staging slots, each with its own stream; chunk k runs H2D -> pair kernel -> D2H
// on slot k % NSLOT, so the two copy engines and the SMs overlap across chunks.
#include <emmintrin.h>

#include <new>
#include <thread>

#include "ofd_common.cuh"

struct ofd_pair_pipeline {
    static constexpr int NSLOT = 3;
    int device, H, W, chunk;
    cudaStream_t st[NSLOT];
    float* d_in[NSLOT];   // img0 (3) | depth0 (1) per frame, frames contiguous per plane group
    float* d_out[NSLOT];  // img1 (3) | depth1 (1) | back_flow (2) | flow (2) | valid (1) | collision (1)
    float* d_s[NSLOT];
    unsigned char* d_u8[NSLOT];  // compact transport staging: img0 u8 (3) in | img1 u8 (3) + valid (1) + collision (1) out
};

namespace ofd {
// compact host transport (ofd_pair_pipeline_run_u8): colour planes and masks cross PCIe as bytes
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}
__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (unsigned char)__float2uint_rn(fminf(fmaxf(in[i], 0.0f), 255.0f));
}
}  // namespace ofd

using namespace ofd;

// constant plane written by the host: non-temporal stores (the plane is not read back by this thread; no read-for-ownership)
static void fill_plane(float* dst, size_t n, float value) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) dst[i++] = value;
    const __m128 v = _mm_set1_ps(value);
    for (; i + 4 <= n; i += 4) _mm_stream_ps(dst + i, v);
    for (; i < n; ++i) dst[i] = value;
    _mm_sfence();
}

#define OFD_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return fail((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

void ofd_pair_pipeline_destroy(ofd_pair_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) {
        if (p->st[s]) cudaStreamSynchronize(p->st[s]), cudaStreamDestroy(p->st[s]);
        cudaFree(p->d_in[s]);
        cudaFree(p->d_out[s]);
        cudaFree(p->d_s[s]);
        cudaFree(p->d_u8[s]);
    }
    delete p;
}

int ofd_pair_pipeline_create(int device, int H, int W, int chunk_frames, ofd_pair_pipeline** out) {
    if (!out) return fail(OFD_E_NULL, "ofd_pair_pipeline_create: out is NULL");
    if (H <= 0 || W <= 0 || chunk_frames <= 0 || chunk_frames > 65535)
        return fail(OFD_E_SHAPE, "ofd_pair_pipeline_create: bad H/W/chunk_frames");
    OFD_CUDA(cudaSetDevice(device));
    ofd_pair_pipeline* p = new (std::nothrow) ofd_pair_pipeline();
    if (!p) return fail(OFD_E_ARG, "ofd_pair_pipeline_create: out of host memory");
    p->device = device, p->H = H, p->W = W, p->chunk = chunk_frames;
    const size_t hw = (size_t)H * W, n = (size_t)chunk_frames;
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) p->st[s] = nullptr, p->d_in[s] = p->d_out[s] = p->d_s[s] = nullptr, p->d_u8[s] = nullptr;
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) {
        cudaError_t e = cudaStreamCreateWithFlags(&p->st[s], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_in[s], n * 4 * hw * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_out[s], n * 10 * hw * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_s[s], n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_u8[s], n * 8 * hw);
        if (e != cudaSuccess) {
            int rc = fail((int)e, "ofd_pair_pipeline_create: %s", cudaGetErrorString(e));
            ofd_pair_pipeline_destroy(p);
            return rc;
        }
    }
    *out = p;
    return OFD_OK;
}

// All *_host pointers are HOST memory (page-locked for overlap), dense [B,C,H,W]; flow/collision may be NULL.
int ofd_pair_pipeline_run(ofd_pair_pipeline* p, const float* img0_host, const float* depth0_host, const float* sBf_host,
                          int B, float* img1_host, float* depth1_host, float* back_flow_host, float* flow_host,
                          float* valid_host, float* collision_host) {
    const char* fn = "ofd_pair_pipeline_run";
    if (!p) return fail(OFD_E_NULL, "%s: pipeline is NULL", fn);
    if (B < 0) return fail(OFD_E_SHAPE, "%s: negative B", fn);
    if (B == 0) return OFD_OK;
    if (!img0_host || !depth0_host || !sBf_host || !img1_host || !depth1_host || !back_flow_host || !valid_host)
        return fail(OFD_E_NULL, "%s: NULL host pointer", fn);
    OFD_CUDA(cudaSetDevice(p->device));
    const size_t hw = (size_t)p->H * p->W, F = sizeof(float);
    // The y planes of both flows are constants of the virtual-stereo pair (flow.y == -0.0, back_flow.y == +0.0, preprocess.py:253,
    // 361-363): they are not sent over PCIe (8 of the 40 result bytes per pixel) but written into the host buffers by a host
    // thread while the copies run.
    constexpr int kFillers = 4;
    std::thread fillers[kFillers];
    struct Joiner {
        std::thread* t;
        ~Joiner() {
            for (int i = 0; i < kFillers; ++i)
                if (t[i].joinable()) t[i].join();
        }
    } joiner{fillers};
    for (int f = 0; f < kFillers; ++f)
        fillers[f] = std::thread([=]() {
            for (int b = f; b < B; b += kFillers) {
                fill_plane(back_flow_host + ((size_t)b * 2 + 1) * hw, hw, 0.0f);
                if (flow_host) fill_plane(flow_host + ((size_t)b * 2 + 1) * hw, hw, -0.0f);
            }
        });
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += p->chunk, ++k) {
        const int s = k % ofd_pair_pipeline::NSLOT;
        const size_t n = (size_t)((B - b0) < p->chunk ? (B - b0) : p->chunk);
        cudaStream_t st = p->st[s];
        float* din = p->d_in[s];
        float* dimg = din;
        float* ddep = din + n * 3 * hw;
        float* dout = p->d_out[s];
        float* o_img = dout;
        float* o_dep = o_img + n * 3 * hw;
        float* o_bf = o_dep + n * hw;
        float* o_fl = o_bf + n * 2 * hw;
        float* o_val = o_fl + n * 2 * hw;
        float* o_col = o_val + n * hw;
        OFD_CUDA(cudaMemcpyAsync(dimg, img0_host + (size_t)b0 * 3 * hw, n * 3 * hw * F, cudaMemcpyHostToDevice, st));
        OFD_CUDA(cudaMemcpyAsync(ddep, depth0_host + (size_t)b0 * hw, n * hw * F, cudaMemcpyHostToDevice, st));
        OFD_CUDA(cudaMemcpyAsync(p->d_s[s], sBf_host + b0, n * F, cudaMemcpyHostToDevice, st));
        int rc = ofd_disparity_pair(dimg, ddep, OFD_F32, p->d_s[s], (int)n, p->H, p->W, o_img, o_dep, o_bf,
                                    flow_host ? o_fl : nullptr, o_val, collision_host ? o_col : nullptr, nullptr, st);
        if (rc) return rc;
        OFD_CUDA(cudaMemcpyAsync(img1_host + (size_t)b0 * 3 * hw, o_img, n * 3 * hw * F, cudaMemcpyDeviceToHost, st));
        OFD_CUDA(cudaMemcpyAsync(depth1_host + (size_t)b0 * hw, o_dep, n * hw * F, cudaMemcpyDeviceToHost, st));
        // x planes only: plane 0 of every [2,H,W] frame (pitch 2*hw floats on both sides)
        OFD_CUDA(cudaMemcpy2DAsync(back_flow_host + (size_t)b0 * 2 * hw, 2 * hw * F, o_bf, 2 * hw * F, hw * F, n, cudaMemcpyDeviceToHost, st));
        if (flow_host)
            OFD_CUDA(cudaMemcpy2DAsync(flow_host + (size_t)b0 * 2 * hw, 2 * hw * F, o_fl, 2 * hw * F, hw * F, n, cudaMemcpyDeviceToHost, st));
        OFD_CUDA(cudaMemcpyAsync(valid_host + (size_t)b0 * hw, o_val, n * hw * F, cudaMemcpyDeviceToHost, st));
        if (collision_host)
            OFD_CUDA(cudaMemcpyAsync(collision_host + (size_t)b0 * hw, o_col, n * hw * F, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) OFD_CUDA(cudaStreamSynchronize(p->st[s]));
    return OFD_OK;  // ~Joiner waits for the host fills
}

// Compact transport of the same pipeline: colour and masks as uint8, the two constant planes (flow.y == -0.0,
// back_flow.y == +0.0) not transferred at all.  Lossless when img0 holds integers 0..255, which is what the reference's
// loader delivers (cv2.imread -> .type(float32), utils.py:17-25); 7 B/px up and 16-17 B/px down instead of 16 and 40.
int ofd_pair_pipeline_run_u8(ofd_pair_pipeline* p, const unsigned char* img0_u8_host, const float* depth0_host,
                             const float* sBf_host, int B, unsigned char* img1_u8_host, float* depth1_host,
                             float* back_flow_x_host, float* flow_x_host, unsigned char* valid_u8_host,
                             unsigned char* collision_u8_host) {
    const char* fn = "ofd_pair_pipeline_run_u8";
    if (!p) return fail(OFD_E_NULL, "%s: pipeline is NULL", fn);
    if (B < 0) return fail(OFD_E_SHAPE, "%s: negative B", fn);
    if (B == 0) return OFD_OK;
    if (!img0_u8_host || !depth0_host || !sBf_host || !img1_u8_host || !depth1_host || !back_flow_x_host || !valid_u8_host)
        return fail(OFD_E_NULL, "%s: NULL host pointer", fn);
    OFD_CUDA(cudaSetDevice(p->device));
    const size_t hw = (size_t)p->H * p->W, F = sizeof(float);
    int k = 0;
    for (int b0 = 0; b0 < B; b0 += p->chunk, ++k) {
        const int s = k % ofd_pair_pipeline::NSLOT;
        const size_t n = (size_t)((B - b0) < p->chunk ? (B - b0) : p->chunk);
        cudaStream_t st = p->st[s];
        float* dimg = p->d_in[s];
        float* ddep = dimg + n * 3 * hw;
        float* o_img = p->d_out[s];
        float* o_dep = o_img + n * 3 * hw;
        float* o_bf = o_dep + n * hw;
        float* o_fl = o_bf + n * 2 * hw;
        float* o_val = o_fl + n * 2 * hw;
        float* o_col = o_val + n * hw;
        unsigned char* u_in = p->d_u8[s];
        unsigned char* u_img = u_in + n * 3 * hw;
        unsigned char* u_val = u_img + n * 3 * hw;
        unsigned char* u_col = u_val + n * hw;
        OFD_CUDA(cudaMemcpyAsync(u_in, img0_u8_host + (size_t)b0 * 3 * hw, n * 3 * hw, cudaMemcpyHostToDevice, st));
        OFD_CUDA(cudaMemcpyAsync(ddep, depth0_host + (size_t)b0 * hw, n * hw * F, cudaMemcpyHostToDevice, st));
        OFD_CUDA(cudaMemcpyAsync(p->d_s[s], sBf_host + b0, n * F, cudaMemcpyHostToDevice, st));
        u8_to_f32_kernel<<<592, 256, 0, st>>>(u_in, dimg, n * 3 * hw);
        int rc = ofd_disparity_pair(dimg, ddep, OFD_F32, p->d_s[s], (int)n, p->H, p->W, o_img, o_dep, o_bf, o_fl, o_val,
                                    collision_u8_host ? o_col : nullptr, nullptr, st);
        if (rc) return rc;
        f32_to_u8_kernel<<<592, 256, 0, st>>>(o_img, u_img, n * 3 * hw);
        f32_to_u8_kernel<<<296, 256, 0, st>>>(o_val, u_val, n * hw);
        if (collision_u8_host) f32_to_u8_kernel<<<296, 256, 0, st>>>(o_col, u_col, n * hw);
        rc = check_launch(fn);
        if (rc) return rc;
        OFD_CUDA(cudaMemcpyAsync(img1_u8_host + (size_t)b0 * 3 * hw, u_img, n * 3 * hw, cudaMemcpyDeviceToHost, st));
        OFD_CUDA(cudaMemcpyAsync(depth1_host + (size_t)b0 * hw, o_dep, n * hw * F, cudaMemcpyDeviceToHost, st));
        // x planes only: plane 0 of every [2,H,W] frame (pitch 2*hw floats)
        OFD_CUDA(cudaMemcpy2DAsync(back_flow_x_host + (size_t)b0 * hw, hw * F, o_bf, 2 * hw * F, hw * F, n, cudaMemcpyDeviceToHost, st));
        if (flow_x_host)
            OFD_CUDA(cudaMemcpy2DAsync(flow_x_host + (size_t)b0 * hw, hw * F, o_fl, 2 * hw * F, hw * F, n, cudaMemcpyDeviceToHost, st));
        OFD_CUDA(cudaMemcpyAsync(valid_u8_host + (size_t)b0 * hw, u_val, n * hw, cudaMemcpyDeviceToHost, st));
        if (collision_u8_host)
            OFD_CUDA(cudaMemcpyAsync(collision_u8_host + (size_t)b0 * hw, u_col, n * hw, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) OFD_CUDA(cudaStreamSynchronize(p->st[s]));
    return OFD_OK;
}

}  // extern "C"
