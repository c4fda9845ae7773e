// ofd_abi.cu — error plumbing, version, workspace management of libofd_b200 (see include/ofd_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>

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
