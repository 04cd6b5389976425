// common.cuh - error plumbing shared by every translation unit of libatvs.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/atvs.h"

void atvs_set_error(const char* fmt, ...);

#define ATVS_CHECK_ARG(cond, code, ...)                   \
    do {                                                  \
        if (!(cond)) {                                    \
            atvs_set_error(__VA_ARGS__);                  \
            return (code);                                \
        }                                                 \
    } while (0)

#define ATVS_CUDA(call)                                                               \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            atvs_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,               \
                           cudaGetErrorString(e__));                                  \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

void atvs_count_launch();
// device counter (one per device, allocated on first use - never during stream capture: every path is warmed up
// eagerly first) of fp16 raw-output rows that had to be clamped to +-65504; read with atvs_saturation_count()
unsigned long long* atvs_sat_ptr();
int atvs_concurrency();          // api.cu: passes the caller runs side by side (atvs_set_concurrency)
#define ATVS_LAUNCH_CHECK()              \
    do {                                 \
        atvs_count_launch();             \
        ATVS_CUDA(cudaPeekAtLastError()); \
    } while (0)

static inline int atvs_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
