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

// ------------------------------------------------------------------------------------------
// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's
// roofline numbers).  Off by default; when on, every launch site brackets its kernel with an
// event pair.  sp_profile_collect() synchronises the events and returns ms / launches per class.
#include <vector>
static bool g_prof_on = false;
struct ProfPair { cudaEvent_t a, b; int cls; };
static std::vector<ProfPair> g_prof_pairs;
static double g_prof_ms[SP_PROF_CLASSES];
static long long g_prof_n[SP_PROF_CLASSES];

void sp_prof_begin(int cls, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfPair p; p.cls = cls;
    cudaEventCreate(&p.a); cudaEventCreate(&p.b);
    cudaEventRecord(p.a, st);
    g_prof_pairs.push_back(p);
}
void sp_prof_end(cudaStream_t st) {
    if (!g_prof_on || g_prof_pairs.empty()) return;
    cudaEventRecord(g_prof_pairs.back().b, st);
}
extern "C" int sp_profile_enable(int on) {
    g_prof_on = on != 0;
    for (int c = 0; c < SP_PROF_CLASSES; c++) { g_prof_ms[c] = 0.0; g_prof_n[c] = 0; }
    for (auto &p : g_prof_pairs) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    g_prof_pairs.clear();
    return SP_OK;
}
extern "C" int sp_profile_collect(double *ms_host, long long *launches_host) {
    for (auto &p : g_prof_pairs) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(p.b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, p.a, p.b);
        if (e != cudaSuccess) return sp_check_cuda(e, "sp_profile_collect");
        g_prof_ms[p.cls] += ms; g_prof_n[p.cls] += 1;
        cudaEventDestroy(p.a); cudaEventDestroy(p.b);
    }
    g_prof_pairs.clear();
    for (int c = 0; c < SP_PROF_CLASSES; c++) {
        if (ms_host) ms_host[c] = g_prof_ms[c];
        if (launches_host) launches_host[c] = g_prof_n[c];
    }
    return SP_OK;
}
