"""torch-CPU fp32 restatement of the 2-D feature extraction module (FEM) ``ResNetDS2SPP``:
/root/reference/cnn_wrapper/atvsnet.py:254-292 (graph), /root/reference/cnn_wrapper/network.py:173-215 (conv_bn on
4-D tensors), :142-170 (conv), :552-616 (bottleneck / res_block), :650-671 (image_resize, avg_pool), :690-693 (concat).

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for SURVEY.md section 8(f) row N1, written ahead of the CUDA
path.  Pinned against the reference's own graph code executed on tests/golden/tf_shim.py
(tests/golden/make_golden_fem.py -> reference_golden_fem.npz); the variable names are the ones that code asks for
(tests/golden/fem_variables.json).  Leaf semantics restated here:
  * tf.layers.conv2d 'SAME' (asymmetric when stride 2: (0,1) on even extents), dilation ``rate`` -> pad (rate, rate);
  * conv_bn: batch-statistics BN (eps 1e-3, biased variance, no affine) then ReLU (network.py:196-215);
  * slim.batch_norm defaults: batch statistics, ``+ beta``, no gamma, eps 1e-3; slim.conv2d: ``+ biases``, ReLU unless
    activation_fn=None; bottleneck conv2 with stride > 1: explicit pad (1,1) then 'VALID' (network.py:589-595);
  * shortcut: identity | [:, ::s, ::s] (1x1 max-pool) | 1x1 conv (+bias, linear) of the PRE-ACTIVATED input;
  * tf.layers.average_pooling2d 'SAME' averages over the valid (unpadded) elements; windows may exceed the map;
  * tf.image.resize_images bilinear, align_corners=True (image_resize always ends up bilinear, network.py:652-656).
"""
import numpy as np
import torch
import torch.nn.functional as TF

F32 = np.float32
BN_EPS = 1e-3


def _same(n, k_eff, s):
    out = -(-n // s)
    total = max((out - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


def conv2d(x, kernel, stride=1, rate=1, padding='SAME', bias=None):
    """x (B,H,W,Cin), kernel [kh,kw,Cin,Cout] -> (B,Ho,Wo,Cout)."""
    B, H, W, _ = x.shape
    kh, kw = kernel.shape[0], kernel.shape[1]
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=F32)).permute(0, 3, 1, 2)
    if padding == 'SAME':
        ph, pw = _same(H, (kh - 1) * rate + 1, stride), _same(W, (kw - 1) * rate + 1, stride)
        t = TF.pad(t, (pw[0], pw[1], ph[0], ph[1]))
    w = torch.from_numpy(np.ascontiguousarray(kernel, dtype=F32)).permute(3, 2, 0, 1).contiguous()
    b = None if bias is None else torch.from_numpy(np.ascontiguousarray(bias, dtype=F32))
    y = TF.conv2d(t, w, b, stride=stride, dilation=rate)
    return np.ascontiguousarray(y.permute(0, 2, 3, 1).numpy())


def batch_norm_train(x, beta=None):
    x = np.asarray(x, dtype=F32)
    axes = tuple(range(x.ndim - 1))
    mean = x.mean(axis=axes, dtype=F32)
    var = np.mean(np.square(x - mean), axis=axes, dtype=F32)
    inv = (F32(1.0) / np.sqrt(var + F32(BN_EPS))).astype(F32)
    y = x * inv - mean * inv
    return y if beta is None else y + np.asarray(beta, dtype=F32)


def relu(x):
    return np.maximum(x, F32(0))


def conv_bn(x, w, name, stride=1, rate=1):
    """network.py:173-215 on a 4-D tensor: name/conv2d/kernel, BN without affine, ReLU."""
    return relu(batch_norm_train(conv2d(x, w[name + '/conv2d/kernel'], stride, rate)))


def bottleneck(x, w, scope, depth, stride=1, rate=1):
    """network.py:552-603 (pre-activation bottleneck, depth_bottleneck == depth)."""
    depth_in = x.shape[-1]
    preact = relu(batch_norm_train(x, w[scope + '/preact/beta']))
    if depth == depth_in:
        shortcut = x if stride == 1 else x[:, ::stride, ::stride, :]
    else:
        shortcut = conv2d(preact, w[scope + '/shortcut/weights'], stride, 1, 'SAME', w[scope + '/shortcut/biases'])
    r = relu(conv2d(preact, w[scope + '/conv1/weights'], 1, 1, 'SAME', w[scope + '/conv1/biases']))
    if stride == 1:
        r = relu(conv2d(r, w[scope + '/conv2/weights'], 1, rate, 'SAME', w[scope + '/conv2/biases']))
    else:
        k_eff = 3 + 2 * (rate - 1)
        beg = (k_eff - 1) // 2
        end = k_eff - 1 - beg
        r = np.pad(r, ((0, 0), (beg, end), (beg, end), (0, 0)))
        r = relu(conv2d(r, w[scope + '/conv2/weights'], stride, rate, 'VALID', w[scope + '/conv2/biases']))
    r = conv2d(r, w[scope + '/conv3/weights'], 1, 1, 'SAME', w[scope + '/conv3/biases'])
    return shortcut + r


def res_block(x, w, name, depth, num_block, stride=1, rate=1):
    """network.py:605-616: blocks name_0 .. name_{n-2}, the last one takes the bare name."""
    if num_block == 1:
        return bottleneck(x, w, name, depth, stride, rate)
    out = bottleneck(x, w, name + '_0', depth, stride, rate)
    for i in range(1, num_block):
        out = bottleneck(out, w, name + '_%d' % i if i != num_block - 1 else name, depth, 1, rate)
    return out


def avg_pool_same(x, k, s):
    B, H, W, C = x.shape
    (pt, _), (pl, _) = _same(H, k, s), _same(W, k, s)
    Ho, Wo = -(-H // s), -(-W // s)
    out = np.empty((B, Ho, Wo, C), F32)
    for i in range(Ho):
        y0, y1 = max(i * s - pt, 0), min(i * s - pt + k, H)
        for j in range(Wo):
            x0, x1 = max(j * s - pl, 0), min(j * s - pl + k, W)
            out[:, i, j] = x[:, y0:y1, x0:x1].sum(axis=(1, 2), dtype=F32) / F32((y1 - y0) * (x1 - x0))
    return out


def resize_bilinear_align(x, Ho, Wo):
    B, H, W, C = x.shape

    def axis(n_in, n_out):
        scale = F32(n_in - 1) / F32(n_out - 1) if n_out > 1 else F32(0)
        src = np.arange(n_out, dtype=F32) * scale
        lo = np.floor(src).astype(np.int64)
        return lo, np.minimum(lo + 1, n_in - 1), (src - lo.astype(F32)).astype(F32)

    y0, y1, fy = axis(H, Ho)
    x0, x1, fx = axis(W, Wo)
    fy, fx = fy[None, :, None, None], fx[None, None, :, None]
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * fx
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * fx
    return (top + (bot - top) * fy).astype(F32)


def ResNetDS2SPP(image, w, return_layers=False):
    """cnn_wrapper/atvsnet.py:254-292: image (B,H,W,3) -> feature (B,H/4,W/4,32)."""
    L = {}
    x = conv_bn(np.asarray(image, dtype=F32), w, 'conv0_0', stride=2)
    x = conv_bn(x, w, 'conv0_1')
    x = L['conv0_2'] = conv_bn(x, w, 'conv0_2')
    x = L['conv0_x'] = res_block(x, w, 'conv0_x', 32, 3, 1, 1)
    c1 = L['conv1_x'] = res_block(x, w, 'conv1_x', 64, 8, 2, 1)
    x = L['conv2_x'] = res_block(c1, w, 'conv2_x', 128, 3, 1, 2)
    c3 = L['conv3_x'] = res_block(x, w, 'conv3_x', 128, 3, 1, 4)
    h, wd = c3.shape[1], c3.shape[2]
    branches = []
    for i, k in enumerate((64, 32, 16, 8)):
        p = avg_pool_same(c3, k, k)
        p = conv_bn(p, w, 'branch_%d_conv' % i)
        branches.append(resize_bilinear_align(p, h, wd))
        L['branch_%d' % i] = branches[-1]
    cat = np.concatenate([c1, c3] + branches, axis=-1)
    f0 = L['fusion0'] = conv_bn(cat, w, 'fusion0')
    out = L['fusion1'] = conv2d(f0, w['fusion1/kernel'], 1, 1)
    return (out, L) if return_layers else out
