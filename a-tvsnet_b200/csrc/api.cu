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
