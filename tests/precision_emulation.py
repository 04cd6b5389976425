"""CPU emulation of the 16-bit tensor-core path's ROUNDING POINTS on top of the fp32 oracle (test infrastructure).

The CUDA path stores: the source feature map and the warped half of the cost volume in the activation dtype, the
convolution weights in the activation dtype, the raw (pre-BN) convolution outputs in fp16 or fp32 (moments always from
the fp32 accumulators), the normalised / ReLU'd / skip-joined activations in the activation dtype, the attention logits
raw, the two 8->1 output convolutions' results in fp32.  This module replays the oracle's stage I + II schedule
(/root/reference/atvsnet/example.py:144-158) with exactly those roundings inserted, for any choice of activation dtype
('bf16' | 'f16' | 'f32'), so that the depth-MAE budget of a storage format can be measured without a GPU and pinned
by a CPU test.  It says nothing about kernel correctness: the -m gpu tests do that against the plain oracle.

    python tests/precision_emulation.py [D h w nviews gain]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om            # noqa: E402
from oracle import network as onet        # noqa: E402
from oracle.homography_warping import get_homographies, homography_warping   # noqa: E402

F32 = np.float32


def quantiser(kind):
    if kind in (None, 'f32'):
        return lambda x: np.asarray(x, dtype=F32)
    tdt = {'bf16': torch.bfloat16, 'f16': torch.float16}[kind]

    def q(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=F32))
        if tdt == torch.float16:
            t = t.clamp(-65504.0, 65504.0)            # cvt.rn.satfinite
        return t.to(tdt).to(torch.float32).numpy()
    return q


class Emu(object):
    """rounding configuration: act = activations + weights, raw = pre-BN conv outputs, skip_sum = dtype of the
    skip-join sums (the BN pass adds in fp32 and rounds once)."""

    def __init__(self, act='bf16', raw='f16', weights=None, overrides=None):
        self.qa = quantiser(act)
        self.qr = quantiser(raw)
        self.qw = quantiser(weights or act)
        self.overrides = overrides or {}      # layer name -> act kind for that layer's OUTPUT

    def qa_for(self, name):
        if name in self.overrides:
            return quantiser(self.overrides[name])
        return self.qa


class EmuUNet(onet.StackedUNet_prob):
    """StackedUNet_prob with the CUDA path's storage roundings."""

    def __init__(self, inputs, weights, emu, first_layer_bias=None):
        self.emu = emu
        self.bias = first_layer_bias or {}
        super(EmuUNet, self).__init__(inputs, weights)

    def _bn_store(self, name, raw):
        e = self.emu
        raw = np.asarray(raw, dtype=F32)
        axes = tuple(range(raw.ndim - 1))
        mean = raw.mean(axis=axes, dtype=np.float64)
        var = np.square(raw.astype(np.float64) - mean).mean(axis=axes)
        inv = (1.0 / np.sqrt(var + onet.BN_EPS)).astype(F32)
        mean = mean.astype(F32)
        rq = e.qr(raw)
        y = np.maximum((rq - mean) * inv, F32(0))
        return y          # fp32, un-rounded: rounded at the store (plain and/or sum)

    def conv_bn(self, k, filters, stride, name):
        e = self.emu
        x = self.terminals[0]
        raw = onet.conv3d(x, e.qw(self.weights[name + '/conv3d/kernel']), stride)
        if name in self.bias:
            raw = raw + self.bias[name]
        y = self._bn_store(name, raw)
        self.layers[name + '#f32'] = y
        return self._done(name, e.qa_for(name)(y))

    def deconv_bn(self, k, filters, stride, name):
        e = self.emu
        raw = onet.deconv3d(self.terminals[0], e.qw(self.weights[name + '/conv3d_transpose/kernel']), stride)
        y = self._bn_store(name, raw)
        self.layers[name + '#f32'] = y
        return self._done(name, e.qa_for(name)(y))

    def feed(self, *names):
        self._fed = names
        return super(EmuUNet, self).feed(*names)

    def add(self, name):
        # the fused pass adds the lead layer's fp32 normalised value to the stored (rounded) skips and rounds once
        lead = self._fed[0]
        out = self.layers.get(lead + '#f32', self.layers[lead])
        for n in self._fed[1:]:
            out = out + self.layers[n]
        return self._done(name, self.emu.qa_for(name)(out))

    def conv(self, k, filters, stride, name, relu=True):
        e = self.emu
        out = onet.conv3d(self.terminals[0], e.qw(self.weights[name + '/kernel']), stride)
        return self._done(name, out)


def warp_half_arith(src16, H):
    """bilinear warp of a (1,h,w,C) feature map (values already on the fp16 grid) with the BLEND evaluated in fp16
    arithmetic, as a packed-half (HFMA2) K1 would: weights rounded to fp16, out = fma(wd, d, fma(wc, c, fma(wb, b, wa*a)))
    with every step rounded to fp16 (numpy float16 rounds after the multiply AND after the add: an upper bound on the
    single rounding of a hardware fma).  Coordinates / validity exactly as oracle.homography_warping (fp32)."""
    f = np.float32
    _, h, w, C = src16.shape
    Hm = np.asarray(H, dtype=f)[0]
    xs, ys = np.meshgrid(np.arange(w, dtype=f) + f(0.5), np.arange(h, dtype=f) + f(0.5))
    xa = (Hm[0, 0] * xs + Hm[0, 1] * ys) + Hm[0, 2]
    ya = (Hm[1, 0] * xs + Hm[1, 1] * ys) + Hm[1, 2]
    z = (Hm[2, 0] * xs + Hm[2, 1] * ys) + Hm[2, 2]
    z = z + f(1e-7) * (z == 0)
    with np.errstate(all='ignore'):
        u = xa / z - f(0.5)
        v = ya / z - f(0.5)
        valid = (u >= 0) & (v >= 0) & (u < f(w - 1)) & (v < f(h - 1)) & ~np.isnan(u) & ~np.isnan(v)
    u = np.where(valid, u, 0).astype(f)
    v = np.where(valid, v, 0).astype(f)
    x0 = np.floor(u).astype(np.int64)
    y0 = np.floor(v).astype(np.int64)
    x1, y1 = np.minimum(x0 + 1, w - 1), np.minimum(y0 + 1, h - 1)
    fx1, fy1 = (x0 + 1).astype(f), (y0 + 1).astype(f)
    wa = ((fy1 - v) * (fx1 - u)).astype(np.float16)[..., None]
    wb = ((fy1 - v) * (u - x0.astype(f))).astype(np.float16)[..., None]
    wc = ((v - y0.astype(f)) * (fx1 - u)).astype(np.float16)[..., None]
    wd = ((v - y0.astype(f)) * (u - x0.astype(f))).astype(np.float16)[..., None]
    img = src16[0].astype(np.float16)
    a, b, c, d = img[y0, x0], img[y0, x1], img[y1, x0], img[y1, x1]
    out = wd * d + (wc * c + (wb * b + wa * a))
    out = np.where(valid[..., None], out, np.float16(0))
    return out.astype(f)[None]


def emu_stage12(features, cams, depth_num, weights, emu, split_first_layer=True, k1_half_arith=False):
    """oracle.model.run_multiview_stage12 (siamese=False) with the storage roundings of ``emu``."""
    cams = np.asarray(cams, dtype=F32)
    features = np.asarray(features, dtype=F32)
    B, N = cams.shape[:2]
    ds, di = cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1]
    filt = []
    for v in range(1, N):
        hs = get_homographies(cams[:, 0], cams[:, v], depth_num, ds, di, True)
        src = emu.qa(features[:, v])
        if k1_half_arith:
            warped = np.stack([warp_half_arith(src, hs[:, d])[0] for d in range(depth_num)], axis=0)[None]
        else:
            warped = np.stack([homography_warping(src, hs[:, d]) for d in range(depth_num)], axis=1)
        warped = emu.qa(warped)
        ref = emu.qa(features[:, 0])
        cv = np.concatenate([np.tile(ref[:, None], (1, depth_num, 1, 1, 1)), warped], axis=-1)
        tower = EmuUNet({'data': cv}, weights, emu)
        filt.append(tower.get_output_by_name('conv_b2_6_1'))
    # AAM: 8->16 conv on the stored filtered volumes, raw logits stored in the raw dtype, softmax in fp32
    wu = emu.qw(weights['attention_aggregate/attention_activation/weight_unique'])
    ws = emu.qw(weights['attention_aggregate/attention_activation/weight_shared'])
    u = [np.maximum(emu.qr(onet.conv3d(x, wu)), 0) for x in filt]
    s = [np.maximum(emu.qr(onet.conv3d(x, ws)), 0) for x in filt]
    ssum = s[0]
    for t in s[1:]:
        ssum = ssum + t
    act = np.stack([(u[n] - s[n]) + ssum for n in range(len(filt))], axis=-1)
    m = act.max(axis=-1, keepdims=True)
    ex = np.exp(act - m)
    score = ex / ex.sum(axis=-1, keepdims=True, dtype=F32)
    cost_agg = (score * np.stack(filt, axis=-1)).sum(axis=-1, dtype=F32)
    prob_agg = onet.conv3d(emu.qa(cost_agg), emu.qw(weights['attention_prob_vol/kernel']))[..., 0]
    depth, depth_up = om.prob2depth_upsample(prob_agg, depth_num, ds, di)
    return dict(depth=depth, depth_up=depth_up, prob_volume_agg=prob_agg, cost_volume_agg=cost_agg)


def depth_mae_over_range(a, b, cams, depth_num):
    rng_ = (depth_num - 1) * float(cams[0, 0, 1, 3, 1])
    return float(np.abs(a - b).mean()) / rng_


def main(argv):
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, 'a-tvsnet_b200', name + '.py'))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    syn, var = load('synthetic'), load('variables')
    D, h, w, nv = (int(a) for a in argv[:4]) if len(argv) >= 4 else (32, 32, 48, 3)
    gain = float(argv[4]) if len(argv) > 4 else 4.0
    cams = syn.orbit_cams(nv, h, w, D)[None]
    feats = syn.smooth_features(nv, h, w, 32, seed=3)[None]
    weights = var.synthetic_weights(seed=11, logit_gain=gain)
    ref = om.run_multiview_stage12(feats, cams, D, weights, siamese=False)
    p = torch.softmax(-torch.from_numpy(ref['prob_volume_agg']), dim=1).max(dim=1).values.mean().item()
    print("D=%d h=%d w=%d views=%d gain=%g  mean peak prob %.3f" % (D, h, w, nv, gain, p))
    for act, raw in (('f32', 'f32'), ('bf16', 'f16'), ('bf16', 'f32'), ('f16', 'f16'), ('f16', 'f32')):
        out = emu_stage12(feats, cams, D, weights, Emu(act, raw))
        print("  act=%-4s raw=%-3s  depth_up MAE/range = %.4e" % (
            act, raw, depth_mae_over_range(out['depth_up'], ref['depth_agg_init_up'], cams, D)), flush=True)
    out = emu_stage12(feats, cams, D, weights, Emu('f16', 'f16'), k1_half_arith=True)
    print("  act=f16  raw=f16 + K1 blend in fp16 arithmetic  depth_up MAE/range = %.4e" % (
        depth_mae_over_range(out['depth_up'], ref['depth_agg_init_up'], cams, D)), flush=True)


if __name__ == '__main__':
    main(sys.argv[1:])
