// fusion.cu - depth-map fusion: the consistency kernel of /root/reference/fusibile
// (fusibile.cu:138-277) and its point-cloud collection (fusibile.cu:279-325, 425-430) on sm_100a.
//
// What is different from the reference's sm_61 kernel, not in the results:
//   * no textures and no managed memory: the (normal, depth) and colour maps are plain (N,H,W,4) fp32 arrays read with
//     coalesced 16-byte loads; the texture unit's linear filter is restated in arithmetic (un-normalised coordinates,
//     clamp, 8-bit fractional weights - CUDA programming guide, "Texture Fetching"), so results do not depend on it;
//   * ALL reference cameras in one launch (grid z = camera) instead of one launch + cudaDeviceSynchronize + host scan of a
//     managed buffer per camera (fusibile.cu:425-430);
//   * the host scan is an ordered device-side compaction (flags -> per-block counts -> scan -> scatter) that keeps the
//     reference's order: camera-major, row-major inside a camera, points with a zero coordinate dropped (:306).
// fp32 with the op order of oracle/fusibile.py (this file is compiled --fmad=false, no fast-math).
#include "common.cuh"

namespace {

constexpr int FU_MAXCAM = 64;

struct FuCam {
    float P[12];
    float Minv[9];
    float C[3];
    float f;
};

struct FuParams {
    int N, H, W;
    float depth_thresh, normal_thresh;
    int num_consistent, save_texture;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// linear-filtered fetch at texture coordinate (x + 0.5, y + 0.5) of an (H,W,4) map
__device__ __forceinline__ float4 tex_linear(const float* __restrict__ img, int H, int W, float x, float y) {
    const float fi = floorf(x), fj = floorf(y);
    const float a = floorf((x - fi) * 256.0f + 0.5f) / 256.0f;
    const float b = floorf((y - fj) * 256.0f + 0.5f) / 256.0f;
    const int i = (int)fi, j = (int)fj;
    const int i0 = min(max(i, 0), W - 1), i1 = min(max(i + 1, 0), W - 1);
    const int j0 = min(max(j, 0), H - 1), j1 = min(max(j + 1, 0), H - 1);
    const float4 t00 = ld4(img + ((size_t)j0 * W + i0) * 4), t01 = ld4(img + ((size_t)j0 * W + i1) * 4);
    const float4 t10 = ld4(img + ((size_t)j1 * W + i0) * 4), t11 = ld4(img + ((size_t)j1 * W + i1) * 4);
    const float w00 = (1.0f - a) * (1.0f - b), w01 = a * (1.0f - b), w10 = (1.0f - a) * b, w11 = a * b;
    float4 r;
    r.x = ((w00 * t00.x + w01 * t01.x) + w10 * t10.x) + w11 * t11.x;
    r.y = ((w00 * t00.y + w01 * t01.y) + w10 * t10.y) + w11 * t11.y;
    r.z = ((w00 * t00.z + w01 * t01.z) + w10 * t10.z) + w11 * t11.z;
    r.w = ((w00 * t00.w + w01 * t01.w) + w10 * t10.w) + w11 * t11.w;
    return r;
}

// one thread per (reference camera, pixel): flag + dense point record
__global__ void __launch_bounds__(256)
k_fuse(const float* __restrict__ nd, const float* __restrict__ img, const FuCam* __restrict__ cams, const FuParams p,
       unsigned char* __restrict__ flags, float4* __restrict__ coord, float4* __restrict__ normal_out,
       float4* __restrict__ tex_out) {
    __shared__ FuCam sc[FU_MAXCAM > 16 ? 16 : FU_MAXCAM];
    const int ref = blockIdx.z;
    const long long hw = (long long)p.H * p.W;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool small = p.N <= 16;
    if (small) {
        for (int i = threadIdx.x; i < p.N * (int)(sizeof(FuCam) / 4); i += blockDim.x)
            reinterpret_cast<float*>(sc)[i] = reinterpret_cast<const float*>(cams)[i];
        __syncthreads();
    }
    const FuCam* cam = small ? sc : cams;
    if (pix >= hw) return;
    const int px = (int)(pix % p.W), py = (int)(pix / p.W);
    const float* ndr = nd + (size_t)ref * hw * 4;
    const float4 nrm = ld4(ndr + pix * 4);
    const FuCam& cr = cam[ref];
    const float depth = nrm.w;
    // get3Dpoint_cu (fusibile.cu:58-67)
    const float ptx = depth * (float)px - cr.P[3], pty = depth * (float)py - cr.P[7], ptz = depth - cr.P[11];
    float X[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) X[r] = (cr.Minv[3 * r] * ptx + cr.Minv[3 * r + 1] * pty) + cr.Minv[3 * r + 2] * ptz;
    float4 cn = nrm;
    float4 ct = make_float4(0.f, 0.f, 0.f, 0.f);
    if (img != nullptr) ct = ld4(img + ((size_t)ref * hw + pix) * 4);
    int count = 0;
    for (int i = 0; i < p.N; ++i) {
        if (i == ref) continue;
        const FuCam& c = cam[i];
        // project_on_camera (fusibile.cu:127-133)
        const float tx = ((c.P[0] * X[0] + c.P[1] * X[1]) + c.P[2] * X[2]) + c.P[3];
        const float ty = ((c.P[4] * X[0] + c.P[5] * X[1]) + c.P[6] * X[2]) + c.P[7];
        const float tz = ((c.P[8] * X[0] + c.P[9] * X[1]) + c.P[10] * X[2]) + c.P[11];
        const float u = tx / tz, v = ty / tz;
        if (!(u >= 0.f && u < (float)p.W && v >= 0.f && v < (float)p.H)) continue;
        const float4 tnd = tex_linear(nd + (size_t)i * hw * 4, p.H, p.W, u, v);
        const float dx = cr.C[0] - c.C[0], dy = cr.C[1] - c.C[1], dz = cr.C[2] - c.C[2];
        const float baseline = sqrtf((dx * dx + dy * dy) + dz * dz);
        const float dd = cr.f * baseline / tz;
        const float td = cr.f * baseline / tnd.w;
        if (!((fabsf(dd - td) / dd) < p.depth_thresh)) continue;
        float ang = acosf((tnd.x * nrm.x + tnd.y * nrm.y) + tnd.z * nrm.z);
        if (ang != ang) ang = 0.f;
        if (!(ang < p.normal_thresh)) continue;
        cn.x += tnd.x; cn.y += tnd.y; cn.z += tnd.z; cn.w += tnd.w;
        if (p.save_texture && img != nullptr) {
            const float4 tt = tex_linear(img + (size_t)i * hw * 4, p.H, p.W, u, v);
            ct.x += tt.x; ct.y += tt.y; ct.z += tt.z; ct.w += tt.w;
        }
        ++count;
    }
    const float div = (float)count + 1.0f;
    const size_t o = (size_t)ref * hw + pix;
    const bool keep = count >= p.num_consistent && X[0] != 0.f && X[1] != 0.f && X[2] != 0.f;
    flags[o] = keep ? 1 : 0;
    if (keep) {
        coord[o] = make_float4(X[0], X[1], X[2], (float)count);
        normal_out[o] = make_float4(cn.x / div, cn.y / div, cn.z / div, cn.w / div);
        if (tex_out != nullptr) tex_out[o] = make_float4(ct.x / div, ct.y / div, ct.z / div, ct.w / div);
    }
}

constexpr int CP_BLOCK = 1024;      // elements per compaction block (256 threads x 4)

__global__ void __launch_bounds__(256) k_count(const unsigned char* __restrict__ flags, long long n, unsigned* __restrict__ counts) {
    const long long base = (long long)blockIdx.x * CP_BLOCK + threadIdx.x * 4;
    int c = 0;
    for (int k = 0; k < 4; ++k)
        if (base + k < n) c += flags[base + k];
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        counts[blockIdx.x] = (unsigned)t;
    }
}

// exclusive scan of the per-block counts by ONE block (a few thousand entries); total -> *total
__global__ void __launch_bounds__(1024) k_scan(unsigned* __restrict__ counts, int nb, long long* __restrict__ total) {
    __shared__ unsigned long long carry;
    __shared__ unsigned s[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nb ? counts[i] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const unsigned t = threadIdx.x >= off ? s[threadIdx.x - off] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        const unsigned incl = s[threadIdx.x];
        const unsigned long long c0 = carry;
        __syncthreads();
        if (i < nb) counts[i] = (unsigned)(c0 + incl - v);
        if (threadIdx.x == 1023) carry = c0 + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (long long)carry;
}

__global__ void __launch_bounds__(256)
k_scatter(const unsigned char* __restrict__ flags, long long n, const unsigned* __restrict__ offsets,
          const float4* __restrict__ coord, const float4* __restrict__ normal_in, const float4* __restrict__ tex_in,
          long long capacity, float* __restrict__ out_coord, float* __restrict__ out_normal, float* __restrict__ out_tex) {
    const long long base = (long long)blockIdx.x * CP_BLOCK + threadIdx.x * 4;
    int f[4], c = 0;
    for (int k = 0; k < 4; ++k) {
        f[k] = (base + k < n) ? flags[base + k] : 0;
        c += f[k];
    }
    // ordered offsets inside the block: warp scan + warp totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = c;
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    __shared__ int ws[8];
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += ws[w];
    long long dst = (long long)offsets[blockIdx.x] + wbase + incl - c;
    for (int k = 0; k < 4; ++k) {
        if (!f[k]) continue;
        if (dst < capacity) {
            const float4 x = coord[base + k], nn = normal_in[base + k];
            out_coord[dst * 3] = x.x; out_coord[dst * 3 + 1] = x.y; out_coord[dst * 3 + 2] = x.z;
            out_normal[dst * 3] = nn.x; out_normal[dst * 3 + 1] = nn.y; out_normal[dst * 3 + 2] = nn.z;
            if (out_tex != nullptr && tex_in != nullptr) reinterpret_cast<float4*>(out_tex)[dst] = tex_in[base + k];
        }
        ++dst;
    }
}

inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

// workspace: flags (n) | dense coord, normal, texture (n float4 each) | block counts
extern "C" size_t atvs_fuse_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    const size_t n = (size_t)N * H * W;
    return al256(n) + 3 * al256(n * 16) + al256(((n + CP_BLOCK - 1) / CP_BLOCK) * 4) + al256(sizeof(FuCam) * FU_MAXCAM);
}

extern "C" int atvs_fuse_depth_maps(const float* normals_depths, const float* images, const float* cam_P, const float* cam_Minv,
                                    const float* cam_C, const float* cam_f, int N, int H, int W, float depth_thresh,
                                    float normal_thresh, int num_consistent, int save_texture, void* workspace,
                                    long long capacity, float* out_coord, float* out_normal, float* out_texture,
                                    long long* out_count, atvs_stream_t stream) {
    ATVS_CHECK_ARG(normals_depths && cam_P && cam_Minv && cam_C && cam_f && workspace && out_coord && out_normal && out_count,
                   ATVS_E_NULL, "atvs_fuse_depth_maps: NULL pointer");
    ATVS_CHECK_ARG(N >= 2 && N <= FU_MAXCAM && H > 0 && W > 0 && capacity >= 0, ATVS_E_SHAPE,
                   "atvs_fuse_depth_maps: N=%d (2..%d) H=%d W=%d", N, FU_MAXCAM, H, W);
    ATVS_CHECK_ARG((((uintptr_t)normals_depths | (uintptr_t)images | (uintptr_t)workspace | (uintptr_t)out_texture) & 15) == 0,
                   ATVS_E_SHAPE, "atvs_fuse_depth_maps: maps and workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)N * H * W;
    char* ws = (char*)workspace;
    unsigned char* flags = (unsigned char*)ws;                      ws += al256(n);
    float4* coord = (float4*)ws;                                    ws += al256(n * 16);
    float4* nrm = (float4*)ws;                                      ws += al256(n * 16);
    float4* tex = (float4*)ws;                                      ws += al256(n * 16);
    unsigned* counts = (unsigned*)ws;                               ws += al256(((n + CP_BLOCK - 1) / CP_BLOCK) * 4);
    FuCam* cams = (FuCam*)ws;
    // camera records are assembled on the device from the caller's four arrays (no host staging: pointers are device)
    for (int i = 0; i < N; ++i) {
        ATVS_CUDA(cudaMemcpyAsync(cams[i].P, cam_P + 12 * i, 48, cudaMemcpyDeviceToDevice, st));
        ATVS_CUDA(cudaMemcpyAsync(cams[i].Minv, cam_Minv + 9 * i, 36, cudaMemcpyDeviceToDevice, st));
        ATVS_CUDA(cudaMemcpyAsync(cams[i].C, cam_C + 3 * i, 12, cudaMemcpyDeviceToDevice, st));
        ATVS_CUDA(cudaMemcpyAsync(&cams[i].f, cam_f + i, 4, cudaMemcpyDeviceToDevice, st));
    }
    FuParams p;
    p.N = N; p.H = H; p.W = W;
    p.depth_thresh = depth_thresh; p.normal_thresh = normal_thresh;
    p.num_consistent = num_consistent; p.save_texture = save_texture;
    const long long hw = (long long)H * W;
    dim3 grid((unsigned)((hw + 255) / 256), 1, (unsigned)N);
    k_fuse<<<grid, 256, 0, st>>>(normals_depths, images, cams, p, flags, coord, nrm,
                                 (images != nullptr && out_texture != nullptr) ? tex : nullptr);
    ATVS_LAUNCH_CHECK();
    const int nb = (int)((n + CP_BLOCK - 1) / CP_BLOCK);
    k_count<<<nb, 256, 0, st>>>(flags, (long long)n, counts);
    ATVS_LAUNCH_CHECK();
    k_scan<<<1, 1024, 0, st>>>(counts, nb, out_count);
    ATVS_LAUNCH_CHECK();
    k_scatter<<<nb, 256, 0, st>>>(flags, (long long)n, counts, coord, nrm,
                                  (images != nullptr && out_texture != nullptr) ? tex : nullptr, capacity, out_coord,
                                  out_normal, out_texture);
    ATVS_LAUNCH_CHECK();
    return 0;
}
