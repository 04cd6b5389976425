// conv_deconv_ring.cu - placeholder until the plane-ring transposed convolution lands
#include "common.cuh"
#include "conv_deconv.cuh"
bool deconv_ring_supported(int, int) { return false; }
bool deconv_ring_applicable(int, int, int, int) { return false; }
size_t deconv_ring_weight_bytes(int, int) { return 0; }
int deconv_ring_pack(const float*, int, int, int, void*, cudaStream_t) { return 0; }
int deconv_ring(const void*, int, const void*, int, int, int, int, int, int, float*, int, double*, cudaStream_t) { return ATVS_E_UNSUP; }
