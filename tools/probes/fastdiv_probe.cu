// fastdiv_probe.cu - evidence for ztest_reproject_kernel (csrc/ofd_splat.cu): the hoisted-reciprocal division
//   r = rcp_refined(b) = MUFU.RCP + one Newton step;  q0 = a * r;  q = fma(r, fma(-b, q0, a), q0)
// equals the IEEE quotient __fdiv_rn(a, b) bit for bit on the guarded operand set
//   |b| in [2^-40, 2^40] and |q| in [2^-40, 2^40]          (u = c0 / (z + eps), v = c1 / (z + eps))
//   b an integer in [1, 2^24], |a| in [2^-40, 2^40]         (u / (W - 1), v / (H - 1))
// Build + run on the GPU box:  nvcc -arch=sm_100a -O3 -fmad=false -prec-div=true -o /tmp/fastdiv tools/probes/fastdiv_probe.cu && /tmp/fastdiv
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float rcp_refined(float b) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b));
    return __fmaf_rn(r0, __fmaf_rn(-b, r0, 1.0f), r0);
}
__device__ __forceinline__ float div_with_rcp(float a, float b, float r) {
    const float q0 = __fmul_rn(a, r);
    return __fmaf_rn(r, __fmaf_rn(-b, q0, a), q0);
}
__device__ __forceinline__ uint64_t mix(uint64_t x) {  // splitmix64
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// a float with a uniformly random mantissa and sign, exponent uniform in [elo, ehi]
__device__ __forceinline__ float rnd_float(uint64_t h, int elo, int ehi) {
    const uint32_t man = (uint32_t)h & 0x7FFFFFu, sign = (uint32_t)(h >> 23) & 1u;
    const int e = elo + (int)((h >> 24) % (uint64_t)(ehi - elo + 1));
    return __uint_as_float((sign << 31) | ((uint32_t)(e + 127) << 23) | man);
}

__global__ void probe(uint64_t seed, int mode, unsigned long long* bad, unsigned long long* tested, float* first_bad) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nb = 0, nt = 0;
    for (int it = 0; it < 4096; ++it) {
        const uint64_t h1 = mix(seed + tid * 4096ull + it), h2 = mix(h1);
        float a, b;
        if (mode == 0) {  // quotient first, so that |q| spans its whole guarded range
            b = rnd_float(h1, -40, 39);
            const float q = rnd_float(h2, -40, 39);
            a = q * b * (1.0f + (float)((h2 >> 40) & 0xFFFF) * 1e-7f);
        } else if (mode == 1) {  // b = W - 1: an integer-valued float
            b = (float)(1u + (uint32_t)(h1 % 16777215ull));
            a = rnd_float(h2, -40, 39);
        } else {  // mantissa patterns near the rounding boundaries: b with few / many ones, a = exact multiples +- 1 ulp
            const uint32_t pat[8] = {0x000000u, 0x7FFFFFu, 0x400000u, 0x000001u, 0x7FFFFEu, 0x555555u, 0x2AAAAAu, 0x3FFFFFu};
            b = __uint_as_float(((uint32_t)(127 + (int)(h1 % 60) - 30) << 23) | pat[(h1 >> 8) & 7]);
            const float q = rnd_float(h2, -30, 29);
            const uint32_t ab = __float_as_uint(__fmul_rn(q, b)) + (uint32_t)((h2 >> 50) % 5) - 2u;
            a = __uint_as_float(ab);
        }
        const float r = rcp_refined(b);
        const float q = div_with_rcp(a, b, r);
        const float aq = fabsf(q), ab_ = fabsf(b);
        const bool guarded = ab_ >= 9.094947017729282e-13f && ab_ <= 1099511627776.0f &&
                             (mode == 1 ? (fabsf(a) >= 9.094947017729282e-13f && fabsf(a) <= 1099511627776.0f)
                                        : (aq >= 9.094947017729282e-13f && aq <= 1099511627776.0f));
        if (!guarded) continue;
        ++nt;
        const float ref = __fdiv_rn(a, b);
        if (__float_as_uint(ref) != __float_as_uint(q)) {
            if (!nb && atomicAdd(bad, 0ull) == 0ull) first_bad[0] = a, first_bad[1] = b, first_bad[2] = q, first_bad[3] = ref;
            ++nb;
        }
    }
    atomicAdd(bad, nb);
    atomicAdd(tested, nt);
}

int main() {
    unsigned long long *bad, *tested;
    float* fb;
    cudaMallocManaged(&bad, 8);
    cudaMallocManaged(&tested, 8);
    cudaMallocManaged(&fb, 16);
    int rc = 0;
    for (int mode = 0; mode < 3; ++mode) {
        *bad = 0, *tested = 0;
        for (int rep = 0; rep < 8; ++rep) probe<<<148 * 32, 256>>>(0x1234567ull + 1000003ull * rep + 77ull * mode, mode, bad, tested, fb);
        if (cudaDeviceSynchronize() != cudaSuccess) return printf("CUDA error\n"), 2;
        printf("mode %d: %llu guarded operand pairs, %llu mismatches vs __fdiv_rn", mode, *tested, *bad);
        if (*bad) printf("  first: a=%a b=%a fast=%a ieee=%a", fb[0], fb[1], fb[2], fb[3]), rc = 1;
        printf("\n");
    }
    return rc;
}
