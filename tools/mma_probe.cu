// mma_probe.cu - cost of one tcgen05.mma (M=128, K=16, bf16) as a function of N, operand layout and
// whether A comes from shared memory or TMEM.  Decides how the convolution kernels batch their taps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include "../a-tvsnet_b200/csrc/tc_ptx.cuh"
#include <cstdlib>
#include <vector>

void atvs_set_error(const char*, ...) {}
void atvs_count_launch() {}

__device__ __forceinline__ void mma_ts_acc(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 q, %4, 0;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(leader)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

// mode 0: SS, no-swizzle ring-like layout (8 rows x 16 B core matrices, SBO = 160, LBO = 2896)
// mode 1: SS, no-swizzle dense (SBO = 128, LBO = 2048)
// mode 2: SS, 128B swizzle K-major (row = 128 B; K=16 uses 32 B of it)
// mode 3: A in TMEM
// mode 4: SS ring-like, but A start address advances by 16 B each MMA (tap shifts) over 3 planes
template <int N, int mode>
__global__ void __launch_bounds__(128) k_probe(int iters, int nacc, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (threadIdx.x < 32) {
        const uint32_t leader = elect_one();
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        uint64_t ad, bd;
        if (mode == 0 || mode == 4) ad = make_desc(a0, 2896, 160, 0);
        else if (mode == 1) ad = make_desc(a0, 2048, 128, 0);
        else ad = make_desc(a0, 16, 1024, 2 /*128B swizzle*/);
        if (mode == 2) bd = make_desc(b0, 16, 1024, 2);
        else bd = make_desc(b0, (uint32_t)N * 16, 128, 0);
        long long t0 = clock64();
        const uint32_t d0 = tmem + 256, d1 = tmem + 256 + (nacc == 2 ? (uint32_t)N % 256u : 0u);
        for (int i = 0; i < iters; i += 27) {
#pragma unroll
            for (int j = 0; j < 27; ++j) {
                const uint32_t d = (j & 1) ? d1 : d0;
                uint64_t a = ad;
                if (mode == 4) a = desc_advance(ad, (uint32_t)((j / 9) * 8192 + (((j % 9) / 3) * 10 + (j % 3)) * 16));
                if (mode == 3) mma_ts_acc(d, tmem, bd, idesc, leader);
                else tc_mma_bf16_acc(d, a, bd, idesc, leader);
            }
        }
        tc_commit_leader(&bar, leader);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (leader) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, int mode>
void run(long long* d, const char* name) {
    const int iters = 2700;
    cudaFuncSetAttribute(k_probe<N, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int nacc : {1, 2}) {
        const int grid = 148;
        k_probe<N, mode><<<grid, 128, 200 * 1024>>>(iters, nacc, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); exit(1); }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        printf("%-22s N=%3d nacc=%d grid=%3d : %.1f cycles / MMA\n", name, N, nacc, grid, (double)mx / iters);
    }
}
template <int mode>
void run_mode(long long* d, const char* name) {
    run<16, mode>(d, name); run<32, mode>(d, name); run<48, mode>(d, name); run<64, mode>(d, name);
    run<96, mode>(d, name); run<128, mode>(d, name); run<256, mode>(d, name);
}
int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    run_mode<0>(d, "SS ring-layout");
    run_mode<1>(d, "SS dense no-swizzle");
    run_mode<2>(d, "SS 128B swizzle");
    run_mode<3>(d, "A in TMEM");
    run_mode<4>(d, "SS ring shifted taps");
    return 0;
}
