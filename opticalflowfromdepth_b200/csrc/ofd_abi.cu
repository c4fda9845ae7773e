// ofd_abi.cu — error plumbing, version, workspace management of libofd_b200 (see include/ofd_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "ofd_common.cuh"

namespace ofd {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return OFD_OK;
}

// Launch-plan cache: the dynamic shared-memory opt-in and the occupancy query of a persistent kernel cost
// 10-20 us of driver time per call - nothing at batch 256, most of a batch-1 drop-in call.  They depend only on
// (device, kernel, threads, smem), so they are resolved once and remembered.  The cache is append-only and mutex-protected; it
// holds no tensor state (the "no global mutable state" rule of the ABI is about data, this is memoised driver metadata).
namespace {
struct PlanEntry {
    int dev;
    const void* kern;
    int threads;
    size_t smem;
    int sms, per_sm;
};
struct SmemEntry {
    int dev;
    const void* kern;
    size_t smem;
};
std::mutex g_plan_mu;
std::vector<PlanEntry> g_plans;
std::vector<SmemEntry> g_smem;
}  // namespace

int ensure_dynamic_smem(const char* fn, const void* kern, size_t smem) {
    if (smem <= 48 * 1024) return OFD_OK;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (auto& e : g_smem)
        if (e.dev == dev && e.kern == kern) {
            if (e.smem >= smem) return OFD_OK;
            cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return fail((int)err, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(err));
            e.smem = smem;
            return OFD_OK;
        }
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return fail((int)err, "%s: cudaFuncSetAttribute: %s", fn, cudaGetErrorString(err));
    g_smem.push_back({dev, kern, smem});
    return OFD_OK;
}

int launch_plan(const char* fn, const void* kern, int threads, size_t smem, int* sms, int* per_sm) {
    int rc = ensure_dynamic_smem(fn, kern, smem);
    if (rc) return rc;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_plan_mu);
        for (const auto& e : g_plans)
            if (e.dev == dev && e.kern == kern && e.threads == threads && e.smem == smem) {
                *sms = e.sms, *per_sm = e.per_sm;
                return OFD_OK;
            }
    }
    int n_sm = 0, occ = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    if (err != cudaSuccess || occ < 1 || n_sm < 1) return fail(err ? (int)err : OFD_E_ARG, "%s: occupancy query failed", fn);
    std::lock_guard<std::mutex> lk(g_plan_mu);
    g_plans.push_back({dev, kern, threads, smem, n_sm, occ});
    *sms = n_sm, *per_sm = occ;
    return OFD_OK;
}

}  // namespace ofd

extern "C" {

int ofd_version(void) { return 100; }

const char* ofd_last_error_string(void) { return ofd::g_err; }

size_t ofd_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    size_t n = (size_t)B * (size_t)H * (size_t)W * sizeof(ofd::u64);
    return (n + 255) & ~(size_t)255;
}

int ofd_workspace_reset(void* ws, size_t bytes, ofd_stream_t stream) {
    if (bytes == 0) return OFD_OK;
    if (!ws) return ofd::fail(OFD_E_NULL, "ofd_workspace_reset: ws is NULL");
    if (((uintptr_t)ws & 15) || (bytes & 15))
        return ofd::fail(OFD_E_WORKSPACE, "ofd_workspace_reset: ws/bytes must be 16-byte aligned");
    // all-ones bytes == KEY_UNTOUCHED; a plain memset is the fastest fill and graph-capturable
    cudaError_t e = cudaMemsetAsync(ws, 0xFF, bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return ofd::fail((int)e, "ofd_workspace_reset: %s", cudaGetErrorString(e));
    return OFD_OK;
}

}  // extern "C"
