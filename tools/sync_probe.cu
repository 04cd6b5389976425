// sync_probe.cu - latency of the synchronisation primitives of the conv pipelines, one warp / one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/sync_probe tools/sync_probe.cu
#include "../a-tvsnet_b200/csrc/tc_ptx.cuh"
#include <vector>
void atvs_set_error(const char*, ...) {}
void atvs_count_launch() {}

__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool try_wait_nohint(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(128) k_sync(long long* out) {
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tslot;
    __shared__ __align__(16) uint8_t buf[4096];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const int N = 512;
    if (threadIdx.x == 0) {
        long long t0, t1;
        // 1. try_wait on a completed phase (parity 1 passes on a fresh barrier)
        t0 = clock64();
        for (int i = 0; i < N; ++i) mbar_wait(&bar[0], 1);
        t1 = clock64();
        out[0] = (t1 - t0);
        // 2. arrive + wait on own barrier (phase flips each time)
        t0 = clock64();
        for (int i = 0; i < N; ++i) { mbar_arrive(&bar[1]); mbar_wait(&bar[1], i & 1); }
        t1 = clock64();
        out[1] = (t1 - t0);
        // 3. tcgen05.commit (no MMA pending) + wait
        t0 = clock64();
        for (int i = 0; i < N; ++i) { tc_commit(&bar[2]); mbar_wait(&bar[2], i & 1); }
        t1 = clock64();
        out[2] = (t1 - t0);
        // 4. fence.proxy.async
        t0 = clock64();
        for (int i = 0; i < N; ++i) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        t1 = clock64();
        out[3] = (t1 - t0);
        // 5. tcgen05 fences
        t0 = clock64();
        for (int i = 0; i < N; ++i) { tc_fence_after(); tc_fence_before(); }
        t1 = clock64();
        out[4] = (t1 - t0);
        // 6. cp.async 16B + commit + wait_group 0 (one element)
        t0 = clock64();
        for (int i = 0; i < N; ++i) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(smem_u32(buf)), "l"(out + 64) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        t1 = clock64();
        out[5] = (t1 - t0);
        // 7. commit without waiting (issue cost)
        t0 = clock64();
        for (int i = 0; i < N; ++i) tc_commit(&bar[3]);
        t1 = clock64();
        out[6] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < N; ++i) while (!test_wait(&bar[0], 1)) {}
        t1 = clock64();
        out[9] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < N; ++i) while (!try_wait_nohint(&bar[0], 1)) {}
        t1 = clock64();
        out[10] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < N; ++i) { mbar_arrive(&bar[1]); while (!test_wait(&bar[1], i & 1)) {} }
        t1 = clock64();
        out[11] = (t1 - t0);
        volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(buf + 2048);
        *flag = 0;
        t0 = clock64();
        for (int i = 0; i < N; ++i) { *flag = i + 1; while (*flag != (uint32_t)(i + 1)) {} }
        t1 = clock64();
        out[12] = (t1 - t0);
        out[15] = N;
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        // 8. tcgen05.ld x8 + wait::ld ; 9. tcgen05.st x8 + wait::st   (warp 1 -> lanes 32..63)
        const uint32_t taddr = tmem + ((uint32_t)32 << 16);
        float v[8];
        long long t0 = clock64();
        for (int i = 0; i < N; ++i) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
            tc_wait_ld();
            v[0] = __uint_as_float(r[0]);
        }
        long long t1 = clock64();
        if (threadIdx.x == 32) out[7] = t1 - t0;
        t0 = clock64();
        for (int i = 0; i < N; ++i) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        t1 = clock64();
        if (threadIdx.x == 32) out[8] = t1 - t0 + (long long)(v[0] != v[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 1024);
    cudaMemset(d, 0, 1024);
    for (int rep = 0; rep < 2; ++rep) {
        k_sync<<<1, 128>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"mbar_wait (already complete)", "mbarrier arrive + wait", "tcgen05.commit + wait", "fence.proxy.async",
                           "tcgen05 fence after+before", "cp.async 16B + commit + wait_group 0", "tcgen05.commit (issue only)",
                           "tcgen05.ld x8 + wait::ld", "tcgen05.st x8 + wait::st", "test_wait loop (already complete)",
                           "try_wait without hint (complete)", "arrive + test_wait loop", "volatile smem flag store + load"};
    for (int i = 0; i < 13; ++i) printf("%-40s %8.1f cycles\n", names[i], (double)h[i] / (double)h[15]);
    return 0;
}
