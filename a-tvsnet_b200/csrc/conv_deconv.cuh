// conv_deconv.cuh - interface of the fused 8-class transposed convolution kernel (conv_deconv.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

bool deconv_fused_applicable(int Cin, int Cout);
size_t deconv_fused_weight_bytes(int Cin, int Cout);
int deconv_fused_pack(const float* kernel, int Cin, int Cout, int dtype, void* wimg, cudaStream_t st);
int deconv_fused(const void* x_bf16, int dtype, const void* wimg, int B, int D, int H, int W, int Cin, int Cout, float* raw_out,
                 int raw16, double* stats, cudaStream_t st);

// plane-ring formulation of the same transposed convolution (conv_deconv_ring.cu)
bool deconv_ring_supported(int Cin, int Cout);
bool deconv_ring_applicable(int B, int D, int H, int W);
size_t deconv_ring_weight_bytes(int Cin, int Cout);
int deconv_ring_pack(const float* kernel, int Cin, int Cout, int dtype, void* wimg, cudaStream_t st);
int deconv_ring(const void* x16, int dtype, const void* wimg, int B, int D, int H, int W, int Cin, int Cout, float* raw_out,
                int raw16, double* stats, cudaStream_t st);
