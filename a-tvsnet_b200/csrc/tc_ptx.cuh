// tc_ptx.cuh - inline-PTX wrappers for the Blackwell async machinery used by the tensor-core
// convolution kernels: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (alloc / mma / commit /
// ld / fences), shared-memory matrix descriptors, and the shared epilogue helpers.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr uint32_t TC_SPIN_LIMIT = 1u << 27;   // bounded waits: trap instead of hanging the GPU

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > TC_SPIN_LIMIT) {
            printf("atvs conv_tc: mbarrier timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (K-major operand tile):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) swizzle mode
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

// 32 values per lane -> lane L ends with the sum over the warp of v[L] (31 shuffles)
__device__ __forceinline__ float warp_transpose_reduce32(float* v, int lane) {
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}


__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// epilogue of one 128-row accumulator tile: TMEM -> registers, raw fp32 store of the real columns,
// per-channel sum / sum of squares folded into per-lane running registers (batch-stat BN).
template <int NPAD>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, uint64_t* tempty_bar, int lane, bool valid, float* op,
                                              int ncols, bool vec4, bool want_stats, float* run) {
    float v[NPAD];
#pragma unroll
    for (int c = 0; c < NPAD; c += 16) tc_ld16(taddr + c, v + c);
    tc_wait_ld();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar);
    if (valid) {
        if (vec4) {
#pragma unroll
            for (int c = 0; c < NPAD; c += 4)
                if (c < ncols) *reinterpret_cast<float4*>(op + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < NPAD; ++c)
                if (c < ncols) op[c] = v[c];
        }
    }
    if (want_stats) {
        if (!valid) {
#pragma unroll
            for (int c = 0; c < NPAD; ++c) v[c] = 0.f;
        }
        if (NPAD == 16) {
            float a[32];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                a[c] = v[c];
                a[16 + c] = v[c] * v[c];
            }
            run[0] += warp_transpose_reduce32(a, lane);
        } else {
#pragma unroll
            for (int h = 0; h < NPAD / 32; ++h) {
                float q[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) q[c] = v[h * 32 + c] * v[h * 32 + c];
                run[2 * h + 1] += warp_transpose_reduce32(q, lane);
                run[2 * h] += warp_transpose_reduce32(v + h * 32, lane);
            }
        }
    }
}

// flush the per-lane running statistics: stats[c] += sum, stats[Cout + c] += sum of squares
template <int NPAD>
__device__ __forceinline__ void flush_stats(double* stats, const float* run, int lane, int Cout, int coff, int ncols) {
    if (NPAD == 16) {
        const int c = lane & 15;
        if (c < ncols) atomicAdd(&stats[(lane < 16 ? 0 : Cout) + coff + c], (double)run[0]);
    } else {
#pragma unroll
        for (int h = 0; h < NPAD / 32; ++h) {
            const int c = h * 32 + lane;
            if (c < ncols) {
                atomicAdd(&stats[coff + c], (double)run[2 * h]);
                atomicAdd(&stats[Cout + coff + c], (double)run[2 * h + 1]);
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace
