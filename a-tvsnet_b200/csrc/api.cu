// api.cu - version / error string / device query of the C ABI (include/atvs.h).
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[512] = "";

void atvs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int atvs_version(void) { return 100; }
extern "C" const char* atvs_last_error(void) { return g_err; }
extern "C" int atvs_device_sm_count(void) { return atvs_num_sms(); }

static std::atomic<long long> g_launches{0};
void atvs_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long atvs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static unsigned long long* g_sat[64] = {nullptr};
unsigned long long* atvs_sat_ptr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (g_sat[dev] == nullptr) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        unsigned long long* p = nullptr;
        if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(p, 0, sizeof(unsigned long long));
        (void)cs;
        g_sat[dev] = p;
    }
    return g_sat[dev];
}
extern "C" long long atvs_saturation_count(int reset) {
    unsigned long long* p = atvs_sat_ptr();
    if (p == nullptr) return -1;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, p, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;     // synchronises the device
    if (reset) cudaMemset(p, 0, sizeof(v));
    return (long long)v;
}
