// Error plumbing for the C ABI: int status codes + a thread-local message (SURVEY.md 8b:
// the reference raises Python exceptions; the Python host layer maps these codes back).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "sparsepoly_b200.h"

static thread_local char g_err[512] = "";

void sp_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sp_check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return SP_OK;
    sp_set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return SP_ERR_CUDA;
}

extern "C" const char *sp_last_error(void) { return g_err; }
extern "C" int sp_abi_version(void) { return SP_ABI_VERSION; }
extern "C" int sp_device_count(int *count_host) {
    if (!count_host) { sp_set_error("sp_device_count: null pointer"); return SP_ERR_INVALID; }
    return sp_check_cuda(cudaGetDeviceCount(count_host), "cudaGetDeviceCount");
}
extern "C" int sp_set_device(int device) {
    return sp_check_cuda(cudaSetDevice(device), "cudaSetDevice");
}
