// conv_ring.cuh - interface of the halo-ring stride-1 convolution kernel (conv_ring.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#ifdef __CUDACC__
#define RING_HD __host__ __device__
#else
#define RING_HD
#endif

// K=16 MMA steps per input plane: 9 in-plane taps x Cin/16, or 5 tap pairs for Cin = 8 (the three z taps
// of a step sit side by side in the MMA N dimension)
RING_HD static inline int ring_nsteps(int Cin) { return Cin >= 16 ? 9 * (Cin / 16) : 5; }

int ring_npad(int Cin, int Cout);
size_t ring_weight_bytes(int Cin, int Cout);
int ring_pack(const float* kernel, int Cin, int Cout, int dtype, void* wimg, cudaStream_t st);
bool ring_applicable(int B, int D, int H, int W, int stride, int transposed);
int ring_conv(const void* x16, int dtype, const void* wimg, int B, int D, int H, int W, int Cin, int Cout, float* raw_out,
              int raw16, double* stats, const float* bias, cudaStream_t st);

// stride-2 variant (conv_ring_s2.cu)
bool ring_s2_supported(int Cin, int Cout);
size_t ring_s2_weight_bytes(int Cin, int Cout);
int ring_s2_pack(const float* kernel, int Cin, int Cout, int dtype, void* wimg, cudaStream_t st);
bool ring_s2_applicable(int B, int D, int H, int W, int Cin, int Cout);
int ring_s2_conv(const void* x16, int dtype, const void* wimg, int B, int D, int H, int W, int Cin, int Cout, float* raw_out,
                 int raw16, double* stats, const float* bias, cudaStream_t st);

// attention aggregation in one kernel (conv_attn_ring.cu): all views convolved 8 -> 16 per plane step, combined in the epilogue
bool attn_ring_applicable(int n_views, int D, int H, int W);
int attn_ring(const void* const* views, int n_views, int dtype, const void* wimg, int B, int D, int H, int W, float* out,
              cudaStream_t st);
