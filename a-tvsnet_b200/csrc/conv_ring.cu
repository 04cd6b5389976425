// conv_ring.cu - stride-1 3x3x3 convolution (network.py:173-215 conv_bn, :142 conv) as a
// tcgen05 implicit GEMM whose A operand is read from a shared-memory RING OF HALO PLANES, so
// every input voxel is fetched from L2/HBM ~1.4x instead of 27x (once per filter tap).
//
//   work unit   = a column of output tiles (16 y x 8 x) over a z segment [z0, z0+zlen)
//   ring slot   = one input z plane with halo: [Cin/8 chunks][18 y][10 x][8 channels] bf16
//                 (no-swizzle K-major core-matrix layout: 8 consecutive x = one 8-row core
//                 matrix, next y row = SBO, next 8-channel chunk = LBO), filled by 4 producer
//                 warps with 16-byte cp.async (zero fill outside the volume = 'SAME' padding;
//                 TMA box loads with 16-byte rows measured ~6 cycles per row and were the
//                 bottleneck), published to the async proxy with fence.proxy.async;
//   one tile    = 27 taps x Cin/16 tcgen05.mma (M=128, N=16/32/64, K=16) whose A descriptors are
//                 just shifted start addresses into three consecutive ring planes; a plane is
//                 released (tcgen05.commit -> mbarrier) when the tile that last needs it retires.
//   Cin = 8     : one K=16 step covers two taps of the same plane (LBO = distance of the taps).
//   warps 0-3 producers | warp 4 MMA issuer | warps 5-8 epilogue, double-buffered TMEM.
#include "tc_ptx.cuh"
#include "conv_ring.cuh"
#include <cstring>

namespace {

constexpr int RG_TY = 16, RG_TX = 8, RG_HH = 18, RG_WW = 10;
constexpr int RG_NVOX = RG_HH * RG_WW;              // voxels of one halo plane (180)
constexpr int RG_KCH_PAD = RG_NVOX * 16 + 16;       // pitch of one 8-channel chunk plane; +16 B keeps the
                                                    // 8 chunk stores of a voxel on distinct banks
constexpr int RG_PRODUCERS = 128;
constexpr int RG_THREADS = 288;

struct RingParams {
    int B, D, H, W;
    int Cout, coff, ncols;
    int nXT, nYT, nZS, ZS;
    int nring;
    int wbytes;
    long long nunits;
};

template <int CIN>
struct RingCfg {
    static constexpr int NKC = CIN / 8;
    static constexpr int SLOT_BYTES = (NKC * RG_KCH_PAD + 127) / 128 * 128;
};

struct Unit {
    int b, x0, y0, z0, zlen;
};

__device__ __forceinline__ Unit decode_unit(const RingParams& p, long long u) {
    Unit r;
    const int zs = (int)(u % p.nZS);
    long long t = u / p.nZS;
    r.x0 = (int)(t % p.nXT) * RG_TX;
    t /= p.nXT;
    r.y0 = (int)(t % p.nYT) * RG_TY;
    r.b = (int)(t / p.nYT);
    r.z0 = zs * p.ZS;
    r.zlen = min(p.ZS, p.D - r.z0);
    return r;
}

// small-channel instances are bound by the latency of the producer / MMA / epilogue handshakes, not by
// any pipe: two co-resident CTAs per SM overlap those latencies (shared memory and TMEM both allow it)
template <int CIN, int NPAD>
struct RingOcc {
    static constexpr int MINB = (CIN <= 16 && NPAD == 16) ? 2 : 1;
};

template <int CIN, int NPAD>
__global__ void __launch_bounds__(RG_THREADS, RingOcc<CIN, NPAD>::MINB)
k_conv3d_ring(const __nv_bfloat16* __restrict__ x, const __grid_constant__ RingParams p,
              const uint8_t* __restrict__ wimg, float* __restrict__ out, double* __restrict__ stats,
              const float* __restrict__ bias) {
    using Cfg = RingCfg<CIN>;
    constexpr uint32_t TMEM_COLS = (2 * NPAD < 32) ? 32u : (uint32_t)(2 * NPAD);
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t B_CHUNK = NPAD * 16;             // one 8-channel chunk of a weight tile
    constexpr uint32_t B_STEP = 2 * B_CHUNK;            // one K=16 step

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t* wsm = smem;
    uint8_t* ring = smem + ((p.wbytes + 127) & ~127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.nring * Cfg::SLOT_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.nring;
    uint64_t* tfull = bars + 2 * p.nring;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int R = p.nring;
    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) {
            mbar_init(&full[s], RG_PRODUCERS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tfull[0], 1);
        mbar_init(&tfull[1], 1);
        mbar_init(&tempty[0], 4);
        mbar_init(&tempty[1], 4);
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===================== producers: global -> ring planes (cp.async, 16 B per op) =====================
        const int ptid = threadIdx.x;
        if (ptid == 0) {
            mbar_expect_tx(wbar, (uint32_t)p.wbytes);
            bulk_copy_g2s(wsm, wimg, (uint32_t)p.wbytes, wbar);
        }
        // up to G planes of cp.async in flight per thread; plane q is published (fence.proxy.async +
        // mbarrier arrive) once plane q+G-1 has been issued.  G <= R-2 keeps the ring deadlock-free.
        const int G = (R >= 6) ? 4 : 2;
        uint32_t cnt = 0, published = 0;
        auto publish_upto = [&](uint32_t upto_excl, int keep) {
            // wait until at most `keep` groups are pending, then publish planes [published, upto_excl)
            if (keep >= 3) asm volatile("cp.async.wait_group 3;" ::: "memory");
            else if (keep == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else if (keep == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (; published < upto_excl; ++published) mbar_arrive(&full[published % R]);
        };
        for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
            const Unit un = decode_unit(p, u);
            for (int zi = un.z0 - 1; zi <= un.z0 + un.zlen; ++zi, ++cnt) {
                const uint32_t slot = cnt % R, par = (cnt / R) & 1;
                mbar_wait(&empty[slot], par ^ 1);
                const bool zok = zi >= 0 && zi < p.D;
                const __nv_bfloat16* zbase = x + (((size_t)un.b * p.D + (zok ? zi : 0)) * p.H) * p.W * CIN;
                const uint32_t dst0 = smem_u32(ring + (size_t)slot * Cfg::SLOT_BYTES);
#pragma unroll 4
                for (int i = ptid; i < Cfg::NKC * RG_NVOX; i += RG_PRODUCERS) {
                    const int c = i % Cfg::NKC, v = i / Cfg::NKC;
                    const int yy = v / RG_WW, xx = v - yy * RG_WW;
                    const int gy = un.y0 - 1 + yy, gx = un.x0 - 1 + xx;
                    const bool ok = zok && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
                    const __nv_bfloat16* src = ok ? zbase + ((size_t)gy * p.W + gx) * CIN + c * 8 : x;
                    const uint32_t dst = dst0 + (uint32_t)(c * RG_KCH_PAD + v * 16);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0)
                                 : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (cnt + 1 - published >= (uint32_t)G) publish_upto(cnt + 2 - G, G - 1);
            }
        }
        if (published < cnt) publish_upto(cnt, 0);
    } else if (warp == 4) {
        // ===================== MMA issuer (converged warp, elected lane issues) =====================
        {
            const uint32_t leader = elect_one();
            mbar_wait(wbar, 0);
            tc_fence_after();
            const uint32_t ring_u32 = smem_u32(ring);
            const uint64_t wdesc0 = make_desc(smem_u32(wsm), B_CHUNK, 128, 0);
            uint32_t cnt = 0;
            long long it = 0;
            for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
                const Unit un = decode_unit(p, u);
                for (int t = 0; t < un.zlen; ++t, ++it) {
                    const int acc = (int)(it & 1);
                    mbar_wait(&tempty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                    for (int dz = (t == 0 ? 0 : 2); dz < 3; ++dz) {
                        const uint32_t c = cnt + t + dz;
                        mbar_wait(&full[c % R], (c / R) & 1);
                    }
                    tc_fence_after();
                    const uint32_t dcol = tmem_base + (uint32_t)(acc * NPAD);
                    // descriptors = per-plane base + compile-time offsets (one 64-bit add per operand)
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        const uint32_t pbase = ring_u32 + ((cnt + t + dz) % R) * (uint32_t)Cfg::SLOT_BYTES;
                        if (CIN >= 16) {
                            const uint64_t abase = make_desc(pbase, RG_KCH_PAD, RG_WW * 16, 0);
                            const uint64_t bbase = desc_advance(wdesc0, (uint32_t)(dz * 9 * (CIN / 16)) * B_STEP);
#pragma unroll
                            for (int tp = 0; tp < 9; ++tp) {
#pragma unroll
                                for (int ks = 0; ks < CIN / 16; ++ks) {
                                    const uint64_t ad = desc_advance(abase, (uint32_t)(2 * ks * RG_KCH_PAD + ((tp / 3) * RG_WW + (tp % 3)) * 16));
                                    const uint64_t bd = desc_advance(bbase, (uint32_t)(tp * (CIN / 16) + ks) * B_STEP);
                                    if (dz == 0 && tp == 0 && ks == 0) tc_mma_bf16_first(dcol, ad, bd, IDESC, leader);
                                    else tc_mma_bf16_acc(dcol, ad, bd, IDESC, leader);
                                }
                            }
                        } else {
                            const uint64_t bbase = desc_advance(wdesc0, (uint32_t)(dz * 5) * B_STEP);
#pragma unroll
                            for (int pr = 0; pr < 5; ++pr) {
                                // tap pairs (0,1) (2,3) (4,5) (6,7) and (7*,8): the 9th tap is paired with a
                                // second, zero-weighted read of tap 7 so that every operand byte is real data
                                // (an out-of-plane phantom multiplies uninitialised shared memory by 0 -> NaN)
                                const int ta = (pr < 4) ? 2 * pr : 7, tb = (pr < 4) ? 2 * pr + 1 : 8;
                                const uint32_t offa = (uint32_t)(((ta / 3) * RG_WW + (ta % 3)) * 16);
                                const uint32_t offb = (uint32_t)(((tb / 3) * RG_WW + (tb % 3)) * 16);
                                const uint32_t lbo = offb - offa;
                                // LBO differs per tap pair: fold it into the constant part of the descriptor
                                const uint64_t ad = desc_advance(make_desc(pbase, 0, RG_WW * 16, 0), offa) |
                                                    ((uint64_t)(lbo >> 4) << 16);
                                const uint64_t bd = desc_advance(bbase, (uint32_t)pr * B_STEP);
                                if (dz == 0 && pr == 0) tc_mma_bf16_first(dcol, ad, bd, IDESC, leader);
                                else tc_mma_bf16_acc(dcol, ad, bd, IDESC, leader);
                            }
                        }
                    }
                    tc_commit_leader(&empty[(cnt + t) % R], leader);
                    if (t == un.zlen - 1) {
                        tc_commit_leader(&empty[(cnt + t + 1) % R], leader);
                        tc_commit_leader(&empty[(cnt + t + 2) % R], leader);
                    }
                    tc_commit_leader(&tfull[acc], leader);
                }
                cnt += (uint32_t)un.zlen + 2;
            }
        }
    } else {
        // ===================== epilogue (4 warps = 128 TMEM lanes) =====================
        const int g = warp & 3;
        const int row = g * 32 + lane;
        const int ty = row >> 3, tx = row & 7;
        constexpr int NRED = 2 * NPAD;
        float run[NRED];
#pragma unroll
        for (int i = 0; i < NRED; ++i) run[i] = 0.f;
        const bool vec4 = (p.ncols & 3) == 0 && (p.Cout & 3) == 0;
        long long it = 0;
        for (long long u = blockIdx.x; u < p.nunits; u += gridDim.x) {
            const Unit un = decode_unit(p, u);
            const int y = un.y0 + ty, x = un.x0 + tx;
            const bool valid = y < p.H && x < p.W;
            const size_t obase = valid ? ((((size_t)un.b * p.D + un.z0) * p.H + y) * p.W + x) * p.Cout + p.coff : 0;
            const size_t zstride = (size_t)p.H * p.W * p.Cout;
            for (int t = 0; t < un.zlen; ++t, ++it) {
                const int acc = (int)(it & 1);
                mbar_wait(&tfull[acc], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * NPAD);
                const float* brow = nullptr;
                if (bias != nullptr && valid) {
                    const int z = un.z0 + t;
                    const int zc = (z == 0) ? 0 : (z == p.D - 1 ? 2 : 1);
                    brow = bias + ((((size_t)un.b * 3 + zc) * p.H + y) * p.W + x) * p.Cout + p.coff;
                }
                epilogue_tile<NPAD>(taddr, &tempty[acc], lane, valid, out + obase + (size_t)t * zstride, p.ncols, vec4,
                                    stats != nullptr, run, brow);
            }
        }
        if (stats != nullptr) flush_stats<NPAD>(stats, run, lane, p.Cout, p.coff, p.ncols);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// weight image of one slab: [step][2 chunks][NPAD rows][8 channels] bf16 (see header comment)
__global__ void k_pack_ring(const float* __restrict__ w, int Cin, int Cout, int npad, int nslabs,
                            __nv_bfloat16* __restrict__ out) {
    const int nsteps = ring_nsteps(Cin);
    const int slab = blockIdx.x / nsteps, step = blockIdx.x % nsteps;
    __nv_bfloat16* o = out + ((size_t)slab * nsteps + step) * 2 * npad * 8;
    for (int i = threadIdx.x; i < 2 * npad * 8; i += blockDim.x) {
        const int chunk = i / (npad * 8), n = (i / 8) % npad, e = i % 8;
        int tap, k;
        if (Cin >= 16) {
            tap = step / (Cin / 16);
            k = ((step % (Cin / 16)) * 2 + chunk) * 8 + e;
        } else {
            const int dz = step / 5, pr = step % 5;
            // pairs (0,1) (2,3) (4,5) (6,7) (7*,8): chunk 0 of the last pair is a zero-weighted re-read of tap 7
            const int tp = (pr < 4) ? 2 * pr + chunk : (chunk == 0 ? -1 : 8);
            tap = (tp >= 0) ? dz * 9 + tp : -1;
            k = e;
        }
        const int co = slab * npad + n;
        float val = 0.f;
        if (tap >= 0 && co < Cout) val = w[((size_t)tap * Cin + k) * Cout + co];
        o[i] = __float2bfloat16_rn(val);
    }
}

template <int CIN, int NPAD>
int launch_ring(const __nv_bfloat16* x, const RingParams& p, const uint8_t* wimg, float* out, double* stats,
                const float* bias, size_t smem, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        ATVS_CUDA(cudaFuncSetAttribute(k_conv3d_ring<CIN, NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    k_conv3d_ring<CIN, NPAD><<<grid, RG_THREADS, smem, st>>>(x, p, wimg, out, stats, bias);
    ATVS_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int ring_npad(int Cin, int Cout) {
    int npad = Cout <= 16 ? 16 : Cout <= 32 ? 32 : 64;
    while (npad > 16 && (size_t)ring_nsteps(Cin) * 2 * npad * 16 > 120 * 1024) npad >>= 1;
    return npad;
}

size_t ring_weight_bytes(int Cin, int Cout) {
    const int npad = ring_npad(Cin, Cout);
    const int nslabs = (Cout + npad - 1) / npad;
    return (size_t)nslabs * ring_nsteps(Cin) * 2 * npad * 16;
}

int ring_pack(const float* kernel, int Cin, int Cout, void* wimg, cudaStream_t st) {
    const int npad = ring_npad(Cin, Cout);
    const int nslabs = (Cout + npad - 1) / npad;
    k_pack_ring<<<nslabs * ring_nsteps(Cin), 128, 0, st>>>(kernel, Cin, Cout, npad, nslabs, (__nv_bfloat16*)wimg);
    ATVS_LAUNCH_CHECK();
    return 0;
}

bool ring_applicable(int B, int D, int H, int W, int stride, int transposed) {
    return !transposed && stride == 1 && (long long)D * H * W >= 32768 && H >= 8 && W >= 8;
}

int ring_conv(const void* x_bf16, const void* wimg, int B, int D, int H, int W, int Cin, int Cout, float* raw_out,
              double* stats, const float* bias, cudaStream_t st) {
    const int npad = ring_npad(Cin, Cout);
    const int nslabs = (Cout + npad - 1) / npad;
    const int sms = atvs_num_sms();
    RingParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.D = D; p.H = H; p.W = W; p.Cout = Cout;
    p.nXT = (W + RG_TX - 1) / RG_TX;
    p.nYT = (H + RG_TY - 1) / RG_TY;
    {   // z segment length: minimise waves * (planes per unit)
        const long long cols = (long long)B * p.nXT * p.nYT;
        const long long slots = (long long)sms * ((Cin <= 16 && ring_npad(Cin, Cout) == 16) ? 2 : 1);
        long long best = -1;
        int bz = D;
        for (int zs = (D < 4 ? D : 4); zs <= D; ++zs) {
            const long long units = cols * ((D + zs - 1) / zs);
            const long long cost = ((units + slots - 1) / slots) * (zs + 2);
            if (best < 0 || cost < best) { best = cost; bz = zs; }
        }
        p.ZS = bz;
        p.nZS = (D + bz - 1) / bz;
        p.nunits = cols * p.nZS;
    }
    p.wbytes = ring_nsteps(Cin) * 2 * npad * 16;
    const size_t slot = ((size_t)(Cin / 8) * RG_KCH_PAD + 127) / 128 * 128;
    const int minb = (Cin <= 16 && npad == 16) ? 2 : 1;
    const size_t budget = (minb == 2 ? 100 : 208) * 1024;
    int nring = (int)((budget - (size_t)((p.wbytes + 127) & ~127)) / slot);
    if (nring > 8) nring = 8;
    if (nring < 4) {
        atvs_set_error("atvs_conv3d_bf16(ring): weights do not fit next to 4 ring planes (Cin=%d Cout=%d)", Cin, Cout);
        return ATVS_E_UNSUP;
    }
    p.nring = nring;
    const size_t smem = 128 + ((p.wbytes + 127) & ~127) + (size_t)nring * slot + (2 * nring + 5) * 8 + 16;
    const int grid = (int)(p.nunits < (long long)sms * minb ? p.nunits : (long long)sms * minb);
    for (int slab = 0; slab < nslabs; ++slab) {
        p.coff = slab * npad;
        p.ncols = (Cout - p.coff < npad) ? Cout - p.coff : npad;
        const uint8_t* wi = (const uint8_t*)wimg + (size_t)slab * p.wbytes;
        int rc = 0;
#define RG_CASE(CI, NP) if (Cin == CI && npad == NP) rc = launch_ring<CI, NP>((const __nv_bfloat16*)x_bf16, p, wi, raw_out, stats, bias, smem, grid, st); else
        RG_CASE(8, 16) RG_CASE(8, 32) RG_CASE(8, 64) RG_CASE(16, 16) RG_CASE(16, 32) RG_CASE(16, 64)
        RG_CASE(32, 16) RG_CASE(32, 32) RG_CASE(32, 64) RG_CASE(64, 16) RG_CASE(64, 32)
        {
            atvs_set_error("atvs_conv3d_bf16(ring): no kernel for Cin=%d N=%d", Cin, npad);
            return ATVS_E_UNSUP;
        }
#undef RG_CASE
        if (rc) return rc;
    }
    return 0;
}
