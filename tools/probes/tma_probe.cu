// Minimal probe for the bilateral TMA experiment: one CTA loads one box of a 2-D float tensor through a tensor map.
//   ./tma_probe W H boxW boxH x y     prints the encode status, the CUDA status and a checksum of the tile
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int cells, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(cells * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(su32(tile)),
                     "l"(&tm), "r"(x), "r"(y), "r"(su32(&bar))
                     : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n"
        "D_%=:\n\t}" ::"r"(su32(&bar)),
        "r"(0)
        : "memory");
    for (int e = threadIdx.x; e < cells; e += blockDim.x) out[e] = tile[e];
}

typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int W = atoi(argv[1]), H = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]), x = atoi(argv[5]), y = atoi(argv[6]);
    std::vector<float> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000) + 1.0f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, (size_t)bw * bh * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: err %d query %d ptr %p\n", (int)ge, (int)q, p);
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstr[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = ((enc_t)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d (W %d H %d box %dx%d at %d,%d)\n", (int)r, W, H, bw, bh, x, y);
    if (r != CUDA_SUCCESS) return 0;
    probe<<<1, 128, (size_t)bw * bh * 4 + 128>>>(tm, x, y, bw * bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> t((size_t)bw * bh);
        cudaMemcpy(t.data(), o, t.size() * 4, cudaMemcpyDeviceToHost);
        double s = 0;
        int zeros = 0;
        for (float v : t) s += v, zeros += (v == 0.0f);
        // expected: element (r, c) of the box = tensor[(y + r) * W + x + c] when inside, else 0
        double want = 0;
        for (int r2 = 0; r2 < bh; ++r2)
            for (int c = 0; c < bw; ++c) {
                const int yy = y + r2, xx = x + c;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) want += h[(size_t)yy * W + xx];
            }
        printf("checksum %.1f expected %.1f zeros %d first %.1f\n", s, want, zeros, t[0]);
    }
    return 0;
}
