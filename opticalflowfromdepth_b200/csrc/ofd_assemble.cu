// ofd_assemble.cu — assembling training samples from the planes of a frame group (BASELINE config 4: in-loop sample synthesis).
// The reference's reader slices one pair out of the pre-baked group array (dataloader.py:93-126) and its writer applied one photometric
// function to an image of the pair (preprocess.py:150-163: brightness scale, one-channel shift, grayscale).  On the device both are plane
// operations dst[0..hw) = f(src): a batch of samples is one launch over a table of them instead of some eighty tensor-indexing kernels.
#include "ofd_common.cuh"

namespace ofd {

__device__ __forceinline__ float plane_f(int op, float p, float a, float g, float b) {
    if (op == OFD_PLANE_SCALE) return __fmul_rn(a, p);   // img * scale                      (preprocess.py:152-153)
    if (op == OFD_PLANE_ADD) return __fadd_rn(a, p);     // img[channel] += shift            (:156-158)
    if (op == OFD_PLANE_GRAY)                            // K = 3 dot product, ascending k    (:160-162)
        return __fadd_rn(__fadd_rn(__fmul_rn(a, 0.2989f), __fmul_rn(g, 0.5870f)), __fmul_rn(b, 0.1140f));
    return a;
}

// grid = (blocks over the plane, ops).  VEC: hw % 4 == 0 and every plane 16-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(256) plane_ops_kernel(const ofd_plane_op* __restrict__ ops, size_t hw, size_t gray_stride) {
    const ofd_plane_op o = ops[blockIdx.y];
    const bool gray = o.op == OFD_PLANE_GRAY;
    if (VEC) {
        const size_t n4 = hw / 4, s4 = gray_stride / 4;
        const float4* __restrict__ s = reinterpret_cast<const float4*>(o.src);
        float4* __restrict__ d = reinterpret_cast<float4*>(o.dst);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 a = __ldg(s + i);
            float4 g = a, b = a;
            if (gray) g = __ldg(s + s4 + i), b = __ldg(s + 2 * s4 + i);
            float4 r;
            r.x = plane_f(o.op, o.p, a.x, g.x, b.x), r.y = plane_f(o.op, o.p, a.y, g.y, b.y);
            r.z = plane_f(o.op, o.p, a.z, g.z, b.z), r.w = plane_f(o.op, o.p, a.w, g.w, b.w);
            __stcs(d + i, r);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
            const float a = __ldg(o.src + i);
            float g = a, b = a;
            if (gray) g = __ldg(o.src + gray_stride + i), b = __ldg(o.src + 2 * gray_stride + i);
            o.dst[i] = plane_f(o.op, o.p, a, g, b);
        }
    }
}

}  // namespace ofd

using namespace ofd;

extern "C" int ofd_plane_ops(const ofd_plane_op* ops_dev, int n_ops, size_t hw, size_t gray_stride, int aligned16, ofd_stream_t stream) {
    const char* fn = "ofd_plane_ops";
    if (n_ops < 0) return fail(OFD_E_SHAPE, "%s: negative n_ops", fn);
    if (n_ops == 0 || hw == 0) return OFD_OK;
    if (!ops_dev) return fail(OFD_E_NULL, "%s: NULL table", fn);
    if (((uintptr_t)ops_dev & 7) != 0) return fail(OFD_E_ARG, "%s: the table must be 8-byte aligned", fn);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = aligned16 && hw % 4 == 0 && gray_stride % 4 == 0;
    const size_t work = vec ? hw / 4 : hw;
    // the planes of one table are independent: enough blocks per plane to fill the GPU even with a handful of planes
    unsigned bx = (unsigned)((work + 256 * 4 - 1) / (256 * 4));
    if (bx < 1) bx = 1;
    if (bx > 148 * 8) bx = 148 * 8;
    for (int o0 = 0; o0 < n_ops; o0 += 65535) {
        const int n = n_ops - o0 < 65535 ? n_ops - o0 : 65535;
        if (vec)
            plane_ops_kernel<true><<<dim3(bx, (unsigned)n), 256, 0, st>>>(ops_dev + o0, hw, gray_stride);
        else
            plane_ops_kernel<false><<<dim3(bx, (unsigned)n), 256, 0, st>>>(ops_dev + o0, hw, gray_stride);
    }
    return check_launch(fn);
}
