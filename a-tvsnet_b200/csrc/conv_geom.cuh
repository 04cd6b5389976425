// conv_geom.cuh - index geometry shared by the fp32 (CUDA-core) and bf16 (tcgen05) 3-D
// convolution kernels.  One description covers the three layer kinds of the hot path
// (network.py:173-215 conv_bn stride 1|2, :511-550 deconv_bn stride 2), all kernel 3, TF
// 'SAME' padding:
//
//   out position   o = j * os + p            j in [0, Dj) x [0, Hj) x [0, Wj)
//   input position i = j * s_in + koff[t]    for tap t of nt[dim] taps, kernel index kidx[t]
//
//   conv  stride 1 : os=1 p=0 s_in=1 taps (k, k-1)            k=0..2       (pad 1,1)
//   conv  stride 2 : os=1 p=0 s_in=2 taps (k, k-before)       before = SAME pad (0 for even n)
//   deconv stride 2: one launch per output parity class p in {0,1}^3, os=2, s_in=1,
//                    p=0 -> taps (k=0, off 0), (k=2, off -1);  p=1 -> tap (k=1, off 0)
//                    (out[2i+k] += in[i] * w[k], cropped to [0, 2n)).
#pragma once
#ifdef __CUDACC__
#define ATVS_HD __host__ __device__
#else
#define ATVS_HD
#endif

struct ConvGeom {
    int B, Din, Hin, Win, Cin, Cout;
    int Dj, Hj, Wj;     // iteration space of this launch
    int Do, Ho, Wo;     // full output extent
    int s_in, os;
    int p[3];           // output parity offset (z, y, x)
    int nt[3];          // taps per dim (z, y, x)
    int koff[3][3];     // input offset per tap
    int kidx[3][3];     // kernel index per tap
};

ATVS_HD static inline int same_pad_before(int n, int k, int s) {
    const int out = (n + s - 1) / s;
    int total = (out - 1) * s + k - n;
    if (total < 0) total = 0;
    return total / 2;
}

// conv (transposed == 0): one geometry.  deconv: cls in [0, 8) selects the parity class.
ATVS_HD static inline ConvGeom make_conv_geom(int B, int D, int H, int W, int Cin, int Cout, int stride, int transposed,
                                      int cls) {
    ConvGeom g;
    g.B = B; g.Din = D; g.Hin = H; g.Win = W; g.Cin = Cin; g.Cout = Cout;
    const int n[3] = {D, H, W};
    if (!transposed) {
        g.s_in = stride; g.os = 1;
        int o[3];
        for (int a = 0; a < 3; ++a) {
            o[a] = (n[a] + stride - 1) / stride;
            const int before = same_pad_before(n[a], 3, stride);
            g.p[a] = 0; g.nt[a] = 3;
            for (int k = 0; k < 3; ++k) { g.koff[a][k] = k - before; g.kidx[a][k] = k; }
        }
        g.Dj = g.Do = o[0]; g.Hj = g.Ho = o[1]; g.Wj = g.Wo = o[2];
    } else {
        g.s_in = 1; g.os = 2;
        g.Dj = D; g.Hj = H; g.Wj = W; g.Do = 2 * D; g.Ho = 2 * H; g.Wo = 2 * W;
        for (int a = 0; a < 3; ++a) {
            const int par = (cls >> (2 - a)) & 1;
            g.p[a] = par;
            for (int k = 0; k < 3; ++k) { g.koff[a][k] = 0; g.kidx[a][k] = 0; }
            if (par == 0) { g.nt[a] = 2; g.kidx[a][0] = 0; g.koff[a][0] = 0; g.kidx[a][1] = 2; g.koff[a][1] = -1; }
            else          { g.nt[a] = 1; g.kidx[a][0] = 1; g.koff[a][0] = 0; }
        }
    }
    return g;
}
