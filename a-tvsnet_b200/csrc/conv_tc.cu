// placeholder until the tcgen05 kernel lands
#include "common.cuh"
extern "C" size_t atvs_packed_weight_bytes(int Cin, int Cout, int transposed) { return 16; }
extern "C" int atvs_pack_conv_weights_bf16(const float*, int, int, int, void*, atvs_stream_t) { atvs_set_error("nyi"); return ATVS_E_UNSUP; }
extern "C" int atvs_conv3d_bf16(const void*, const void*, int, int, int, int, int, int, int, int, float*, double*, atvs_stream_t) { atvs_set_error("nyi"); return ATVS_E_UNSUP; }
