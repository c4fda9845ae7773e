// ofd_decode.cu — the input side of the path on the device: JPEG decode and T.Resize.
//
//   utils.get_img (utils.py:17-25): cv2.imread(path, -1) on the host -> BGR uint8 HWC -> float32 CHW.  ReDWeb's images are JPEGs
//   (dataloader.py:28-29): ofd_jpeg_decode hands the compressed bytes to nvJPEG (the toolkit's library, loaded with dlopen so that
//   libofd_b200.so has no link-time dependency on it), which decodes on the GPU into planar B | G | R bytes; a small kernel widens them to
//   the float32 [3,H,W] tensor the pipeline takes.  A frame then crosses PCIe once, as compressed bytes.  JPEG decoding is not bit-specified
//   (IDCT and chroma-upsampling variants): the result is within a few grey levels of libjpeg-turbo's (what cv2 links), measured by the tests.
//   PNG (DIML images, all depth / disparity maps) is deflate - serial, no device decoder in this image: those are decoded on the host and
//   their 8 / 16-bit payload crosses PCIe at 1-2 B/px (ofd_depth_from_png does the loaders' arithmetic on the device).
//
//   T.Resize(img_size)(depth) (dataloader.py:31-32,57-58): torchvision's bilinear resize WITH antialiasing of a [1,H,W] float64 tensor, i.e.
//   ATen's separable upsample_bilinear2d_aa: per output index a triangle filter of support max(scale, 1) around (i + 0.5) * scale, weights
//   normalised to sum 1, horizontal pass then vertical pass, each tap accumulated with a fused multiply-add.  ofd_resize_bilinear_aa restates
//   it: bit-identical to torchvision when upscaling, within 2 ulp when downscaling (ATen's vectorised reduction order differs there).
#include <dlfcn.h>

#include <mutex>

#include "ofd_common.cuh"

// ---- the slice of nvjpeg.h this file needs (declared here so that the build does not depend on the header's include path) -----------
extern "C" {
typedef struct nvjpegHandle* nvjpegHandle_t;
typedef struct nvjpegJpegState* nvjpegJpegState_t;
typedef struct {
    unsigned char* channel[4];
    size_t pitch[4];
} ofd_nvjpegImage_t;
typedef int (*pfn_nvjpegCreateSimple)(nvjpegHandle_t*);
typedef int (*pfn_nvjpegDestroy)(nvjpegHandle_t);
typedef int (*pfn_nvjpegJpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*);
typedef int (*pfn_nvjpegJpegStateDestroy)(nvjpegJpegState_t);
typedef int (*pfn_nvjpegGetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, int*, int*, int*);
typedef int (*pfn_nvjpegDecode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, int, ofd_nvjpegImage_t*, cudaStream_t);
}
constexpr int kNvjpegOutputBGR = 4;  // NVJPEG_OUTPUT_BGR: planar B, G, R

struct ofd_jpeg_decoder {
    int device;
    nvjpegHandle_t handle;
    nvjpegJpegState_t state;
    unsigned char* staging;  // device, grow-only: 3 planes of H*W bytes
    size_t staging_cap;
};

namespace ofd {

struct NvjpegApi {
    void* lib = nullptr;
    pfn_nvjpegCreateSimple create = nullptr;
    pfn_nvjpegDestroy destroy = nullptr;
    pfn_nvjpegJpegStateCreate state_create = nullptr;
    pfn_nvjpegJpegStateDestroy state_destroy = nullptr;
    pfn_nvjpegGetImageInfo info = nullptr;
    pfn_nvjpegDecode decode = nullptr;
};

static const NvjpegApi* nvjpeg_api() {
    static NvjpegApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12",
                               "/usr/local/cuda/targets/x86_64-linux/lib/libnvjpeg.so.12"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.create = (pfn_nvjpegCreateSimple)dlsym(api.lib, "nvjpegCreateSimple");
        api.destroy = (pfn_nvjpegDestroy)dlsym(api.lib, "nvjpegDestroy");
        api.state_create = (pfn_nvjpegJpegStateCreate)dlsym(api.lib, "nvjpegJpegStateCreate");
        api.state_destroy = (pfn_nvjpegJpegStateDestroy)dlsym(api.lib, "nvjpegJpegStateDestroy");
        api.info = (pfn_nvjpegGetImageInfo)dlsym(api.lib, "nvjpegGetImageInfo");
        api.decode = (pfn_nvjpegDecode)dlsym(api.lib, "nvjpegDecode");
        if (!api.create || !api.destroy || !api.state_create || !api.state_destroy || !api.info || !api.decode) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    return api.lib ? &api : nullptr;
}

__global__ void __launch_bounds__(256) bgr_u8_to_f32_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}

// ---- antialiased bilinear resize (ATen upsample_bilinear2d_aa, align_corners = false) ----------------------------------------------------
// One pass along one axis: out[o] = sum_j w_j * in[xmin + j], the triangle filter of ATen's _compute_indices_min_size_weights_aa.
template <typename T>
struct AaTap {
    int xmin, xsize;
    T center, invscale, total;
};

template <typename T>
__device__ __forceinline__ T aa_filter(T x) {
    x = x < (T)0 ? -x : x;
    return x < (T)1 ? (T)1 - x : (T)0;
}

template <typename T>
__device__ __forceinline__ AaTap<T> aa_setup(int o, int in_size, int out_size) {
    const T scale = (T)in_size / (T)out_size;
    const T support = scale >= (T)1 ? ((T)2 * (T)0.5) * scale : (T)2 * (T)0.5;
    AaTap<T> t;
    t.invscale = scale >= (T)1 ? (T)1 / scale : (T)1;
    t.center = scale * ((T)o + (T)0.5);
    int lo = (int)(t.center - support + (T)0.5);
    t.xmin = lo < 0 ? 0 : lo;
    int hi = (int)(t.center + support + (T)0.5);
    hi = hi > in_size ? in_size : hi;
    t.xsize = hi - t.xmin;
    T total = (T)0;
    for (int j = 0; j < t.xsize; ++j) total += aa_filter<T>(((T)(j + t.xmin) - t.center + (T)0.5) * t.invscale);
    t.total = total;
    return t;
}

// src[n_outer, n_in, inner] -> dst[n_outer, n_out, inner] along the middle axis (inner = 1: rows; inner = W: columns of an image)
template <typename T>
__global__ void __launch_bounds__(256) resize_aa_axis_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t n_outer, int n_in, int n_out,
                                                            size_t inner) {
    const size_t total = n_outer * (size_t)n_out * inner;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t in_idx = e % inner;
        const size_t rest = e / inner;
        const int o = (int)(rest % (size_t)n_out);
        const size_t outer = rest / (size_t)n_out;
        const AaTap<T> t = aa_setup<T>(o, n_in, n_out);
        const T* s = src + (outer * (size_t)n_in + (size_t)t.xmin) * inner + in_idx;
        T acc = (T)0;
        for (int j = 0; j < t.xsize; ++j) {
            T w = aa_filter<T>(((T)(j + t.xmin) - t.center + (T)0.5) * t.invscale);
            if (t.total != (T)0) w = w / t.total;
            const T v = s[(size_t)j * inner];
            acc = j == 0 ? v * w : fma(v, w, acc);  // ATen: t = src[0] * w[0]; t += src[j] * w[j] (contracted to fused multiply-adds)
        }
        dst[e] = acc;
    }
}

template <typename T>
static int resize_aa(const char* fn, const T* src, int B, int H, int W, int Ho, int Wo, T* dst, T* tmp, cudaStream_t st) {
    const T* cur = src;
    int blocks;
    if (Wo != W) {  // horizontal pass first (ATen's separable kernel walks the last dimension first)
        T* out = (Ho != H) ? tmp : dst;
        const size_t total = (size_t)B * H * Wo;
        blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        resize_aa_axis_kernel<T><<<blocks, 256, 0, st>>>(cur, out, (size_t)B * H, W, Wo, (size_t)1);
        cur = out;
    }
    if (Ho != H) {
        const size_t total = (size_t)B * Ho * Wo;
        blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        resize_aa_axis_kernel<T><<<blocks, 256, 0, st>>>(cur, dst, (size_t)B, H, Ho, (size_t)Wo);
    } else if (Wo == W) {
        cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)B * H * W * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
    }
    return check_launch(fn);
}

}  // namespace ofd

using namespace ofd;

extern "C" {

int ofd_jpeg_decoder_create(int device, ofd_jpeg_decoder** out) {
    const char* fn = "ofd_jpeg_decoder_create";
    if (!out) return fail(OFD_E_NULL, "%s: out is NULL", fn);
    const NvjpegApi* api = nvjpeg_api();
    if (!api) return fail(OFD_E_ARG, "%s: libnvjpeg.so.12 could not be loaded (%s)", fn, dlerror() ? dlerror() : "symbols missing");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
    ofd_jpeg_decoder* d = new (std::nothrow) ofd_jpeg_decoder();
    if (!d) return fail(OFD_E_ARG, "%s: out of host memory", fn);
    d->device = device, d->handle = nullptr, d->state = nullptr, d->staging = nullptr, d->staging_cap = 0;
    int rc = api->create(&d->handle);
    if (rc == 0) rc = api->state_create(d->handle, &d->state);
    if (rc != 0) {
        if (d->handle) api->destroy(d->handle);
        delete d;
        return fail(OFD_E_ARG, "%s: nvJPEG initialisation failed with status %d", fn, rc);
    }
    *out = d;
    return OFD_OK;
}

void ofd_jpeg_decoder_destroy(ofd_jpeg_decoder* d) {
    if (!d) return;
    const NvjpegApi* api = nvjpeg_api();
    cudaSetDevice(d->device);
    if (api) {
        if (d->state) api->state_destroy(d->state);
        if (d->handle) api->destroy(d->handle);
    }
    cudaFree(d->staging);
    delete d;
}

int ofd_jpeg_info(ofd_jpeg_decoder* d, const uint8_t* data_host, size_t nbytes, int* H, int* W, int* components) {
    const char* fn = "ofd_jpeg_info";
    if (!d || !data_host || !H || !W) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    const NvjpegApi* api = nvjpeg_api();
    int ncomp = 0, sub = 0, ws[4] = {0, 0, 0, 0}, hs[4] = {0, 0, 0, 0};
    const int rc = api->info(d->handle, data_host, nbytes, &ncomp, &sub, ws, hs);
    if (rc != 0) return fail(OFD_E_ARG, "%s: not a JPEG stream nvJPEG can read (status %d)", fn, rc);
    *H = hs[0], *W = ws[0];
    if (components) *components = ncomp;
    return OFD_OK;
}

int ofd_jpeg_decode(ofd_jpeg_decoder* d, const uint8_t* data_host, size_t nbytes, float* out_bgr, int H, int W, ofd_stream_t stream) {
    const char* fn = "ofd_jpeg_decode";
    if (!d || !data_host || !out_bgr) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    const NvjpegApi* api = nvjpeg_api();
    int h = 0, w = 0;
    int rc = ofd_jpeg_info(d, data_host, nbytes, &h, &w, nullptr);
    if (rc) return rc;
    if (h != H || w != W) return fail(OFD_E_SHAPE, "%s: the stream is %dx%d, the output tensor %dx%d", fn, h, w, H, W);
    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
    const size_t hw = (size_t)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->staging_cap < 3 * hw) {
        if (d->staging) {
            cudaStreamSynchronize(st);  // an earlier decode on this stream may still read the old buffer
            cudaFree(d->staging);
        }
        d->staging = nullptr, d->staging_cap = 0;
        e = cudaMalloc(&d->staging, 3 * hw);
        if (e != cudaSuccess) return fail((int)e, "%s: %s", fn, cudaGetErrorString(e));
        d->staging_cap = 3 * hw;
    }
    ofd_nvjpegImage_t img = {};
    for (int c = 0; c < 3; ++c) img.channel[c] = d->staging + (size_t)c * hw, img.pitch[c] = (size_t)W;
    rc = api->decode(d->handle, d->state, data_host, nbytes, kNvjpegOutputBGR, &img, st);
    if (rc != 0) return fail(OFD_E_ARG, "%s: nvjpegDecode failed with status %d", fn, rc);
    bgr_u8_to_f32_kernel<<<296, 256, 0, st>>>(d->staging, out_bgr, 3 * hw);
    return check_launch(fn);
}

int ofd_resize_bilinear_aa(const void* src, int dtype, int B, int H, int W, int H_out, int W_out, void* dst, void* tmp, ofd_stream_t stream) {
    const char* fn = "ofd_resize_bilinear_aa";
    if (dtype != OFD_F32 && dtype != OFD_F64) return fail(OFD_E_DTYPE, "%s: bad dtype %d", fn, dtype);
    if (B < 0 || H <= 0 || W <= 0 || H_out <= 0 || W_out <= 0) return fail(OFD_E_SHAPE, "%s: bad dimension", fn);
    if (B == 0) return OFD_OK;
    if (!src || !dst) return fail(OFD_E_NULL, "%s: NULL tensor pointer", fn);
    if (H_out != H && W_out != W && !tmp) return fail(OFD_E_NULL, "%s: a two-axis resize needs tmp (B * H * W_out elements)", fn);
    if (dtype == OFD_F32) return resize_aa<float>(fn, (const float*)src, B, H, W, H_out, W_out, (float*)dst, (float*)tmp, (cudaStream_t)stream);
    return resize_aa<double>(fn, (const double*)src, B, H, W, H_out, W_out, (double*)dst, (double*)tmp, (cudaStream_t)stream);
}

}  // extern "C"
