// ofd_host.cu — host-buffer front end of the flow-pair path: what a caller holding numpy / pinned CPU tensors
// binds (the reference moves every frame host->device->host around its FW call, preprocess.py:350-366,437-447).
// A pipeline owns NSLOT device staging slots, each with its own stream; chunk k runs H2D -> pair kernel -> D2H
// on slot k % NSLOT, so the two copy engines and the SMs overlap across chunks.
#include <immintrin.h>
#include <sched.h>

#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "ofd_common.cuh"

// One run's work for the host threads: the constant planes to fill and the mask bytes to expand, chunk by chunk.
struct ofd_host_job {
    int B = 0, K = 0, chunk = 0;
    size_t hw = 0;
    bool mask_bytes = false, fill_const = true, img_bytes = false;
    float *back_flow = nullptr, *flow = nullptr, *valid = nullptr, *collision = nullptr, *img1 = nullptr;
    const unsigned char* h_mask = nullptr;
    const unsigned char* h_img8 = nullptr;
    const cudaEvent_t* ev = nullptr;
};

struct ofd_pair_pipeline {
    static constexpr int NSLOT = 3;
    int device, H, W, chunk;
    cudaStream_t st[NSLOT];
    cudaEvent_t done[NSLOT];  // blocking-sync events: run() sleeps on them instead of spinning in cudaStreamSynchronize
    float* d_in[NSLOT];   // img0 (3) | depth0 (1) per frame, frames contiguous per plane group
    float* d_out[NSLOT];  // img1 (3) | depth1 (1) | back_flow (2) | flow (2) | valid (1) | collision (1)
    float* d_s[NSLOT];
    unsigned char* d_u8[NSLOT];  // compact transport staging: img0 u8 (3) in | img1 u8 (3) + valid (1) + collision (1) out
    // float32 contract: valid / collision cross PCIe as ONE byte per pixel (bit 0 / bit 1) into this pinned buffer and are
    // expanded into the caller's float planes by host threads (grow-only, B * H * W bytes)
    unsigned char* h_mask;
    size_t h_mask_cap;
    // img1 is a selection of img0's pixels (or 0): when img0 holds uint8 values - what the reference's loader delivers (cv2.imread ->
    // float32, utils.py:17-25) - img1 does too, and its three planes cross PCIe as bytes (3 instead of 12 B/px) and are widened into the
    // caller's float planes by the host threads.  OPTIMISTIC and VERIFIED: the packing kernel flags every chunk that holds a value a byte
    // cannot carry exactly; such chunks are redone with float planes at the end of the run and the pipeline stops trying (img_bytes_ok).
    unsigned char* h_img8;
    size_t h_img8_cap;
    int* h_flags;
    size_t h_flags_cap;
    int* d_flags[NSLOT];
    bool img_bytes_enabled, img_bytes_ok;
    std::vector<cudaEvent_t> ev;  // one per chunk of a run: "this chunk's mask bytes have landed"
    bool mask_bytes_enabled;
    bool spin_sync;  // default: spinning event waits; OFD_HOST_SYNC=block uses blocking-sync events (sleeping waits, ~4 % slower)
    // Host threads of the pipeline (OFD_HOST_WORKERS, read when the pipeline is created): started once, they sleep on a
    // condition variable between runs and between chunks - no thread creation and no spinning on the timed path.  They
    // inherit the CPU affinity of the thread that created the pipeline (sweep.bind_rank_cores pins a rank to its own cores).
    int n_workers;
    std::vector<std::thread> pool;
    std::mutex mu;
    std::condition_variable cv_job, cv_issued, cv_done;
    unsigned long long job_gen = 0;
    int issued = 0, finished = 0;
    bool abort_run = false, quit = false;
    ofd_host_job job;
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
// float planes -> bytes, four values per thread; *flag is raised when a value is not exactly a uint8 (the chunk is then redone as floats)
__global__ void __launch_bounds__(256) pack_img_u8_kernel(const float4* __restrict__ in, uchar4* __restrict__ out, size_t n4, int* __restrict__ flag) {
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        uchar4 o;
        o.x = (unsigned char)(int)fminf(fmaxf(v.x, 0.0f), 255.0f);
        o.y = (unsigned char)(int)fminf(fmaxf(v.y, 0.0f), 255.0f);
        o.z = (unsigned char)(int)fminf(fmaxf(v.z, 0.0f), 255.0f);
        o.w = (unsigned char)(int)fminf(fmaxf(v.w, 0.0f), 255.0f);
        // exact round trip, bit pattern included (-0.0f would come back as +0.0f)
        bad |= __float_as_uint((float)o.x) != __float_as_uint(v.x) || __float_as_uint((float)o.y) != __float_as_uint(v.y) ||
               __float_as_uint((float)o.z) != __float_as_uint(v.z) || __float_as_uint((float)o.w) != __float_as_uint(v.w);
        out[i] = o;
    }
    if (bad) *flag = 1;
}
// valid (bit 0) and collision (bit 1) of four pixels per thread
__global__ void __launch_bounds__(256) pack_masks_kernel(const float4* __restrict__ valid, const float4* __restrict__ collision,
                                                         uchar4* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = valid[i];
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (collision) c = collision[i];
        uchar4 o;
        o.x = (unsigned char)((v.x != 0.f) | ((c.x != 0.f) << 1));
        o.y = (unsigned char)((v.y != 0.f) | ((c.y != 0.f) << 1));
        o.z = (unsigned char)((v.z != 0.f) | ((c.z != 0.f) << 1));
        o.w = (unsigned char)((v.w != 0.f) | ((c.w != 0.f) << 1));
        out[i] = o;
    }
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

// bit `shift` of 16 mask bytes -> 16 floats (0.0f / 1.0f), non-temporal stores (dst 16-byte aligned)
static inline void expand16(__m128i b, int shift, float* dst) {
    const __m128i z = _mm_setzero_si128();
    const __m128i m = _mm_and_si128(_mm_srl_epi16(b, _mm_cvtsi32_si128(shift)), _mm_set1_epi8(1));
    const __m128i lo = _mm_unpacklo_epi8(m, z), hi = _mm_unpackhi_epi8(m, z);
    const __m128 f0 = _mm_cvtepi32_ps(_mm_unpacklo_epi16(lo, z)), f1 = _mm_cvtepi32_ps(_mm_unpackhi_epi16(lo, z));
    const __m128 f2 = _mm_cvtepi32_ps(_mm_unpacklo_epi16(hi, z)), f3 = _mm_cvtepi32_ps(_mm_unpackhi_epi16(hi, z));
    _mm_stream_ps(dst, f0), _mm_stream_ps(dst + 4, f1), _mm_stream_ps(dst + 8, f2), _mm_stream_ps(dst + 12, f3);
}

static void expand_plane(const unsigned char* m, size_t n, int shift, float* dst) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) dst[i] = (float)((m[i] >> shift) & 1), ++i;
    for (; i + 16 <= n; i += 16) expand16(_mm_loadu_si128((const __m128i*)(m + i)), shift, dst + i);
    for (; i < n; ++i) dst[i] = (float)((m[i] >> shift) & 1);
    _mm_sfence();
}

// bytes -> floats with non-temporal stores.  Two builds of the inner loop: SSE2 (baseline x86-64) and AVX2 (32 bytes per step, picked at
// run time with __builtin_cpu_supports) - one host thread widens ~14 GB/s of floats with the former, about twice that with the latter
static void widen_sse2(const unsigned char* m, size_t n, float* dst) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) dst[i] = (float)m[i], ++i;
    const __m128i z = _mm_setzero_si128();
    for (; i + 16 <= n; i += 16) {
        const __m128i b = _mm_loadu_si128((const __m128i*)(m + i));
        const __m128i lo = _mm_unpacklo_epi8(b, z), hi = _mm_unpackhi_epi8(b, z);
        _mm_stream_ps(dst + i, _mm_cvtepi32_ps(_mm_unpacklo_epi16(lo, z)));
        _mm_stream_ps(dst + i + 4, _mm_cvtepi32_ps(_mm_unpackhi_epi16(lo, z)));
        _mm_stream_ps(dst + i + 8, _mm_cvtepi32_ps(_mm_unpacklo_epi16(hi, z)));
        _mm_stream_ps(dst + i + 12, _mm_cvtepi32_ps(_mm_unpackhi_epi16(hi, z)));
    }
    for (; i < n; ++i) dst[i] = (float)m[i];
    _mm_sfence();
}

__attribute__((target("avx2"))) static void widen_avx2(const unsigned char* m, size_t n, float* dst) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) dst[i] = (float)m[i], ++i;
    for (; i + 32 <= n; i += 32) {
        const __m128i b0 = _mm_loadu_si128((const __m128i*)(m + i)), b1 = _mm_loadu_si128((const __m128i*)(m + i + 16));
        _mm256_stream_ps(dst + i, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(b0)));
        _mm256_stream_ps(dst + i + 8, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(b0, 8))));
        _mm256_stream_ps(dst + i + 16, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(b1)));
        _mm256_stream_ps(dst + i + 24, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(b1, 8))));
    }
    for (; i < n; ++i) dst[i] = (float)m[i];
    _mm_sfence();
}

static void widen_plane(const unsigned char* m, size_t n, float* dst) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2)
        widen_avx2(m, n, dst);
    else
        widen_sse2(m, n, dst);
}

static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    const int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

#define OFD_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return fail((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ---- the pipeline's host threads --------------------------------------------------------------------------------------
static void worker_body(ofd_pair_pipeline* p, int t, const ofd_host_job& J) {
    const int nw = p->n_workers;
    for (int k = 0; k < J.K; ++k) {
        const int b0 = k * J.chunk, n = (J.B - b0) < J.chunk ? (J.B - b0) : J.chunk;
        if (J.fill_const)
            for (int b = b0 + t; b < b0 + n; b += nw) {
                fill_plane(J.back_flow + ((size_t)b * 2 + 1) * J.hw, J.hw, 0.0f);
                if (J.flow) fill_plane(J.flow + ((size_t)b * 2 + 1) * J.hw, J.hw, -0.0f);
            }
        if (!J.mask_bytes) continue;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_issued.wait(lk, [&] { return p->issued > k || p->abort_run; });
            if (p->abort_run) return;
        }
        if (cudaEventSynchronize(J.ev[k]) != cudaSuccess) return;  // the main thread reports the stream's error
        // this worker's 16-pixel-aligned share of the chunk
        const size_t len = (size_t)n * J.hw, per = ((len + nw - 1) / nw + 15) & ~(size_t)15;
        const size_t lo = per * t < len ? per * t : len, hi = lo + per < len ? lo + per : len;
        if (hi > lo) {
            const size_t o = (size_t)b0 * J.hw + lo;
            expand_plane(J.h_mask + o, hi - lo, 0, J.valid + o);
            if (J.collision) expand_plane(J.h_mask + o, hi - lo, 1, J.collision + o);
        }
        if (J.img_bytes) {  // this worker's share of the chunk's 3 * n * hw image bytes
            const size_t len3 = 3 * len, per3 = ((len3 + nw - 1) / nw + 15) & ~(size_t)15;
            const size_t lo3 = per3 * t < len3 ? per3 * t : len3, hi3 = lo3 + per3 < len3 ? lo3 + per3 : len3;
            if (hi3 > lo3) widen_plane(J.h_img8 + 3 * (size_t)b0 * J.hw + lo3, hi3 - lo3, J.img1 + 3 * (size_t)b0 * J.hw + lo3);
        }
    }
}

static void worker_main(ofd_pair_pipeline* p, int t) {
    cudaSetDevice(p->device);
    unsigned long long seen = 0;
    for (;;) {
        ofd_host_job J;
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_job.wait(lk, [&] { return p->quit || p->job_gen != seen; });
            if (p->quit) return;
            seen = p->job_gen;
            J = p->job;
        }
        worker_body(p, t, J);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            ++p->finished;
        }
        p->cv_done.notify_one();
    }
}

// posts a run's host work; the returned guard waits for the workers (and aborts them if the run fails early)
struct HostRun {
    ofd_pair_pipeline* p;
    bool ok = false;
    explicit HostRun(ofd_pair_pipeline* p_, const ofd_host_job& J) : p(p_) {
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->job = J;
            p->issued = 0, p->finished = 0, p->abort_run = false;
            ++p->job_gen;
        }
        p->cv_job.notify_all();
    }
    void chunk_issued(int k) {
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->issued = k + 1;
        }
        p->cv_issued.notify_all();
    }
    ~HostRun() {
        std::unique_lock<std::mutex> lk(p->mu);
        if (!ok) {
            p->abort_run = true;
            p->cv_issued.notify_all();
        }
        p->cv_done.wait(lk, [&] { return p->finished == p->n_workers; });
    }
};

extern "C" {

void ofd_pair_pipeline_destroy(ofd_pair_pipeline* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->quit = true;
    }
    p->cv_job.notify_all();
    for (auto& th : p->pool)
        if (th.joinable()) th.join();
    cudaSetDevice(p->device);
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) {
        if (p->st[s]) cudaStreamSynchronize(p->st[s]), cudaStreamDestroy(p->st[s]);
        if (p->done[s]) cudaEventDestroy(p->done[s]);
        cudaFree(p->d_in[s]);
        cudaFree(p->d_out[s]);
        cudaFree(p->d_s[s]);
        cudaFree(p->d_u8[s]);
    }
    for (cudaEvent_t e : p->ev) cudaEventDestroy(e);
    if (p->h_mask) cudaFreeHost(p->h_mask);
    if (p->h_img8) cudaFreeHost(p->h_img8);
    if (p->h_flags) cudaFreeHost(p->h_flags);
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) cudaFree(p->d_flags[s]);
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
    p->h_mask = nullptr, p->h_mask_cap = 0;
    p->h_img8 = nullptr, p->h_img8_cap = 0, p->h_flags = nullptr, p->h_flags_cap = 0;
    // Bytes for img1 trade 9 B/px of PCIe for 12 B/px of host stores: a win while PCIe is the limit (1-2 GPUs per host: 6.0 -> 6.8 k
    // pairs/s), a loss once the node's host memory is (the D2H total of these boxes is flat from 2 to 8 GPUs).  Default:
    // on unless more than two ranks share the node (torchrun's LOCAL_WORLD_SIZE); OFD_HOST_IMG_BYTES=0/1 overrides.
    p->img_bytes_enabled = env_int("OFD_HOST_IMG_BYTES", env_int("LOCAL_WORLD_SIZE", 1, 1, 1 << 20) <= 2 ? 1 : 0, 0, 1) != 0, p->img_bytes_ok = true;
    // the knobs are read per pipeline (not once per process): a caller can build pipelines with different settings
    // default: one worker per core of this thread's affinity mask minus the issuing thread, between 2 and 4 (the byte-widening of img1
    // wants 3-4: profiles/r2/tune_e2e_img_bytes.txt; under torchrun a rank pinned to 4 cores gets 3)
    int dflt_workers = 2;
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0) {
            const int cores = CPU_COUNT(&set);
            dflt_workers = cores - 1 < 2 ? 2 : (cores - 1 > 4 ? 4 : cores - 1);
        }
    }
    p->n_workers = env_int("OFD_HOST_WORKERS", dflt_workers, 1, 64);
    p->mask_bytes_enabled = env_int("OFD_HOST_MASK_BYTES", 1, 0, 1) != 0;
    {
        // measured on B200 boxes (profiles/r2/tune_e2e.txt): spinning waits 6.35 k pairs/s, blocking-sync events 6.12 k
        const char* e = getenv("OFD_HOST_SYNC");
        p->spin_sync = !(e && (e[0] == 'b' || e[0] == 'B'));
    }
    const size_t hw = (size_t)H * W, n = (size_t)chunk_frames;
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s)
        p->st[s] = nullptr, p->done[s] = nullptr, p->d_in[s] = p->d_out[s] = p->d_s[s] = nullptr, p->d_u8[s] = nullptr, p->d_flags[s] = nullptr;
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) {
        cudaError_t e = cudaStreamCreateWithFlags(&p->st[s], cudaStreamNonBlocking);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&p->done[s], cudaEventDisableTiming | (p->spin_sync ? 0u : (unsigned)cudaEventBlockingSync));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_in[s], n * 4 * hw * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_out[s], n * 10 * hw * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_s[s], n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->d_u8[s], n * 8 * hw);
        if (e == cudaSuccess) e = cudaMalloc(&p->d_flags[s], sizeof(int));
        if (e != cudaSuccess) {
            int rc = fail((int)e, "ofd_pair_pipeline_create: %s", cudaGetErrorString(e));
            ofd_pair_pipeline_destroy(p);
            return rc;
        }
    }
    try {
        for (int t = 0; t < p->n_workers; ++t) p->pool.emplace_back(worker_main, p, t);
    } catch (...) {
        ofd_pair_pipeline_destroy(p);
        return fail(OFD_E_ARG, "ofd_pair_pipeline_create: cannot start host threads");
    }
    *out = p;
    return OFD_OK;
}

// All *_host pointers are HOST memory (page-locked for overlap), dense [B,C,H,W]; flow/collision may be NULL.
int ofd_pair_pipeline_run_flags(ofd_pair_pipeline* p, const float* img0_host, const float* depth0_host, const float* sBf_host,
                                int B, float* img1_host, float* depth1_host, float* back_flow_host, float* flow_host,
                                float* valid_host, float* collision_host, unsigned flags) {
    const char* fn = "ofd_pair_pipeline_run";
    if (!p) return fail(OFD_E_NULL, "%s: pipeline is NULL", fn);
    if (B < 0) return fail(OFD_E_SHAPE, "%s: negative B", fn);
    if (flags & ~(unsigned)OFD_PIPE_KEEP_CONST_PLANES) return fail(OFD_E_ARG, "%s: unknown flag bits 0x%x", fn, flags);
    if (B == 0) return OFD_OK;
    if (!img0_host || !depth0_host || !sBf_host || !img1_host || !depth1_host || !back_flow_host || !valid_host)
        return fail(OFD_E_NULL, "%s: NULL host pointer", fn);
    OFD_CUDA(cudaSetDevice(p->device));
    const size_t hw = (size_t)p->H * p->W, F = sizeof(float);
    // Two things do not cross PCIe as float planes (14 of the 40 result bytes per pixel):
    //  - the y planes of both flows are constants of the virtual-stereo pair (flow.y == -0.0, back_flow.y == +0.0,
    //    preprocess.py:253,361-363) and are written into the host buffers by host threads (OFD_PIPE_KEEP_CONST_PLANES: the
    //    caller recycles result buffers whose y planes already hold the constants - nothing is written there);
    //  - valid / collision are 0.0f / 1.0f planes: they are packed on the device into one byte per pixel, land in the
    //    pipeline's pinned staging buffer and are expanded into the caller's float planes by the same host threads,
    //    chunk by chunk, as soon as a chunk's event fires.   OFD_HOST_MASK_BYTES=0 sends them as float planes instead.
    const bool mask_bytes = p->mask_bytes_enabled && hw % 4 == 0;
    //  - img1 is a selection of img0's pixels: uint8-valued whenever img0 is (the reference's loader output).  Its planes cross as
    //    bytes, verified by the packing kernel chunk by chunk; a chunk a byte cannot carry exactly is redone with float planes below
    //    and the pipeline then stops trying.  OFD_HOST_IMG_BYTES=0 disables it.  Rides on the mask events (needs the byte-mask path).
    const bool img_bytes = mask_bytes && p->img_bytes_enabled && p->img_bytes_ok;
    const int K = (B + p->chunk - 1) / p->chunk;
    if (mask_bytes) {
        if (p->h_mask_cap < (size_t)B * hw) {
            if (p->h_mask) cudaFreeHost(p->h_mask);
            p->h_mask = nullptr, p->h_mask_cap = 0;
            OFD_CUDA(cudaHostAlloc((void**)&p->h_mask, (size_t)B * hw, cudaHostAllocDefault));
            p->h_mask_cap = (size_t)B * hw;
        }
        while ((int)p->ev.size() < K) {
            cudaEvent_t e;
            OFD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | (p->spin_sync ? 0u : (unsigned)cudaEventBlockingSync)));
            p->ev.push_back(e);
        }
    }
    if (img_bytes) {
        if (p->h_img8_cap < (size_t)B * 3 * hw) {
            if (p->h_img8) cudaFreeHost(p->h_img8);
            p->h_img8 = nullptr, p->h_img8_cap = 0;
            OFD_CUDA(cudaHostAlloc((void**)&p->h_img8, (size_t)B * 3 * hw, cudaHostAllocDefault));
            p->h_img8_cap = (size_t)B * 3 * hw;
        }
        if (p->h_flags_cap < (size_t)K) {
            if (p->h_flags) cudaFreeHost(p->h_flags);
            p->h_flags = nullptr, p->h_flags_cap = 0;
            OFD_CUDA(cudaHostAlloc((void**)&p->h_flags, (size_t)K * sizeof(int), cudaHostAllocDefault));
            p->h_flags_cap = (size_t)K;
        }
    }
    {
        ofd_host_job J;
        J.B = B, J.K = K, J.chunk = p->chunk, J.hw = hw;
        J.mask_bytes = mask_bytes, J.fill_const = !(flags & OFD_PIPE_KEEP_CONST_PLANES), J.img_bytes = img_bytes;
        J.back_flow = back_flow_host, J.flow = flow_host, J.valid = valid_host, J.collision = collision_host, J.img1 = img1_host;
        J.h_mask = p->h_mask, J.h_img8 = p->h_img8, J.ev = p->ev.data();
        HostRun host(p, J);
        for (int k = 0, b0 = 0; b0 < B; b0 += p->chunk, ++k) {
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
            if (mask_bytes) {
                unsigned char* u_mask = p->d_u8[s];
                unsigned char* u_img = u_mask + n * hw;
                pack_masks_kernel<<<296, 256, 0, st>>>((const float4*)o_val, collision_host ? (const float4*)o_col : nullptr, (uchar4*)u_mask,
                                                       n * hw / 4);
                if (img_bytes) {
                    OFD_CUDA(cudaMemsetAsync(p->d_flags[s], 0, sizeof(int), st));
                    pack_img_u8_kernel<<<592, 256, 0, st>>>((const float4*)o_img, (uchar4*)u_img, n * 3 * hw / 4, p->d_flags[s]);
                }
                rc = check_launch(fn);
                if (rc) return rc;
                // bytes first: the host expansion of this chunk overlaps the float planes' copies
                OFD_CUDA(cudaMemcpyAsync(p->h_mask + (size_t)b0 * hw, u_mask, n * hw, cudaMemcpyDeviceToHost, st));
                if (img_bytes) {
                    OFD_CUDA(cudaMemcpyAsync(p->h_img8 + (size_t)b0 * 3 * hw, u_img, n * 3 * hw, cudaMemcpyDeviceToHost, st));
                    OFD_CUDA(cudaMemcpyAsync(p->h_flags + k, p->d_flags[s], sizeof(int), cudaMemcpyDeviceToHost, st));
                }
                OFD_CUDA(cudaEventRecord(p->ev[(size_t)k], st));
                host.chunk_issued(k);
            }
            if (!img_bytes)
                OFD_CUDA(cudaMemcpyAsync(img1_host + (size_t)b0 * 3 * hw, o_img, n * 3 * hw * F, cudaMemcpyDeviceToHost, st));
            OFD_CUDA(cudaMemcpyAsync(depth1_host + (size_t)b0 * hw, o_dep, n * hw * F, cudaMemcpyDeviceToHost, st));
            // x planes only: plane 0 of every [2,H,W] frame (pitch 2*hw floats on both sides)
            OFD_CUDA(cudaMemcpy2DAsync(back_flow_host + (size_t)b0 * 2 * hw, 2 * hw * F, o_bf, 2 * hw * F, hw * F, n, cudaMemcpyDeviceToHost, st));
            if (flow_host)
                OFD_CUDA(cudaMemcpy2DAsync(flow_host + (size_t)b0 * 2 * hw, 2 * hw * F, o_fl, 2 * hw * F, hw * F, n, cudaMemcpyDeviceToHost, st));
            if (!mask_bytes) {
                OFD_CUDA(cudaMemcpyAsync(valid_host + (size_t)b0 * hw, o_val, n * hw * F, cudaMemcpyDeviceToHost, st));
                if (collision_host)
                    OFD_CUDA(cudaMemcpyAsync(collision_host + (size_t)b0 * hw, o_col, n * hw * F, cudaMemcpyDeviceToHost, st));
            }
        }
        for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) OFD_CUDA(cudaEventRecord(p->done[s], p->st[s]));
        for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) OFD_CUDA(cudaEventSynchronize(p->done[s]));
        host.ok = true;
    }  // ~HostRun waits for the host fills and expansions
    if (img_bytes) {
        // verification: chunks whose image a byte cannot carry exactly are redone with float planes (img1 only; the rest is in place)
        for (int k = 0, b0 = 0; b0 < B; b0 += p->chunk, ++k) {
            if (!p->h_flags[k]) continue;
            p->img_bytes_ok = false;  // this caller's images are not uint8-valued: later runs send float planes straight away
            const size_t n = (size_t)((B - b0) < p->chunk ? (B - b0) : p->chunk);
            cudaStream_t st = p->st[0];
            float* dimg = p->d_in[0];
            float* ddep = dimg + n * 3 * hw;
            float* o_img = p->d_out[0];
            float* o_dep = o_img + n * 3 * hw;
            float* o_bf = o_dep + n * hw;
            float* o_val = o_bf + 4 * n * hw;
            OFD_CUDA(cudaMemcpyAsync(dimg, img0_host + (size_t)b0 * 3 * hw, n * 3 * hw * F, cudaMemcpyHostToDevice, st));
            OFD_CUDA(cudaMemcpyAsync(ddep, depth0_host + (size_t)b0 * hw, n * hw * F, cudaMemcpyHostToDevice, st));
            OFD_CUDA(cudaMemcpyAsync(p->d_s[0], sBf_host + b0, n * F, cudaMemcpyHostToDevice, st));
            int rc = ofd_disparity_pair(dimg, ddep, OFD_F32, p->d_s[0], (int)n, p->H, p->W, o_img, o_dep, o_bf, nullptr, o_val, nullptr, nullptr, st);
            if (rc) return rc;
            OFD_CUDA(cudaMemcpyAsync(img1_host + (size_t)b0 * 3 * hw, o_img, n * 3 * hw * F, cudaMemcpyDeviceToHost, st));
            OFD_CUDA(cudaStreamSynchronize(st));
        }
    }
    return OFD_OK;
}

int ofd_pair_pipeline_run(ofd_pair_pipeline* p, const float* img0_host, const float* depth0_host, const float* sBf_host,
                          int B, float* img1_host, float* depth1_host, float* back_flow_host, float* flow_host,
                          float* valid_host, float* collision_host) {
    return ofd_pair_pipeline_run_flags(p, img0_host, depth0_host, sBf_host, B, img1_host, depth1_host, back_flow_host, flow_host,
                                       valid_host, collision_host, 0u);
}

int ofd_copy_rows_to_host(const void* src, size_t src_pitch_bytes, void* dst_host, size_t dst_pitch_bytes, size_t width_bytes,
                          size_t rows, ofd_stream_t stream) {
    const char* fn = "ofd_copy_rows_to_host";
    if (!rows || !width_bytes) return OFD_OK;
    if (!src || !dst_host) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    if (src_pitch_bytes < width_bytes || dst_pitch_bytes < width_bytes) return fail(OFD_E_ARG, "%s: pitch smaller than the row width", fn);
    OFD_CUDA(cudaMemcpy2DAsync(dst_host, dst_pitch_bytes, src, src_pitch_bytes, width_bytes, rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return OFD_OK;
}

// float32 -> uint8 with verification (the sweep's group sink sends the 18 image channels of a group as bytes when they are uint8-valued)
__global__ void __launch_bounds__(256) pack_u8_tail_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, size_t n, int* __restrict__ flag) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = in[i];
        const unsigned char o = (unsigned char)(int)fminf(fmaxf(v, 0.0f), 255.0f);
        if (__float_as_uint((float)o) != __float_as_uint(v)) *flag = 1;
        out[i] = o;
    }
}

int ofd_pack_u8(const float* src, uint8_t* dst, size_t n, int* flag, ofd_stream_t stream) {
    const char* fn = "ofd_pack_u8";
    if (!n) return OFD_OK;
    if (!src || !dst || !flag) return fail(OFD_E_NULL, "%s: NULL pointer", fn);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (((uintptr_t)src & 15) == 0) && (((uintptr_t)dst & 3) == 0);
    const size_t n4 = vec ? n / 4 : 0;
    if (n4) pack_img_u8_kernel<<<592, 256, 0, st>>>((const float4*)src, (uchar4*)dst, n4, flag);
    if (n - 4 * n4) pack_u8_tail_kernel<<<n4 ? 1 : 592, 256, 0, st>>>(src + 4 * n4, dst + 4 * n4, n - 4 * n4, flag);
    return check_launch(fn);
}

int ofd_host_widen_u8(const uint8_t* src_host, size_t n, float* dst_host) {
    if (!n) return OFD_OK;
    if (!src_host || !dst_host) return fail(OFD_E_NULL, "ofd_host_widen_u8: NULL pointer");
    widen_plane(src_host, n, dst_host);
    return OFD_OK;
}

int ofd_host_stream_fill(float* dst_host, size_t n, float value) {
    if (!dst_host && n) return fail(OFD_E_NULL, "ofd_host_stream_fill: dst is NULL");
    fill_plane(dst_host, n, value);
    return OFD_OK;
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
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) OFD_CUDA(cudaEventRecord(p->done[s], p->st[s]));
    for (int s = 0; s < ofd_pair_pipeline::NSLOT; ++s) OFD_CUDA(cudaEventSynchronize(p->done[s]));
    return OFD_OK;
}

}  // extern "C"
