"""torch-CPU fp32 restatement of the layer primitives and network graphs used on the hot
path: /root/reference/cnn_wrapper/network.py (conv :142-170, conv_bn :173-215,
attention_activation :282-351, attention_aggregation :379-408, deconv_bn :511-550,
add :696) and /root/reference/cnn_wrapper/atvsnet.py (StackedUNet :5-96,
StackedUNet_prob :100-192, AttAggregation* :196-213, OutputConv* :216-226).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Tensors are NumPy fp32 in the
reference's NDHWC layout; torch-CPU is used only as a conv primitive, always with the
TensorFlow 'SAME' padding made explicit.

Weights come in a dict keyed by the TF variable names of the checkpoint
(SURVEY.md Appendix B), e.g. ``conv_b0_1_0/conv3d/kernel`` [3,3,3,Cin,Cout],
``conv_b0_4_0/conv3d_transpose/kernel`` [3,3,3,Cout,Cin], ``conv_b2_6_2/kernel``.
"""
import numpy as np
import torch
import torch.nn.functional as TF

F32 = np.float32
BN_EPS = 1e-3   # tf.layers.batch_normalization default epsilon (network.py:206)


def _same_pad(n, k, s):
    """TF SAME: out = ceil(n/s); pad_total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _to_t(x):          # NDHWC numpy -> NCDHW torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=F32)).permute(0, 4, 1, 2, 3)


def _to_n(t):          # NCDHW torch -> NDHWC numpy
    return np.ascontiguousarray(t.permute(0, 2, 3, 4, 1).numpy())


def conv3d(x, kernel, stride=1):
    """tf.layers.conv3d / tf.nn.conv3d, padding SAME, no bias.  x (B,D,H,W,Cin),
    kernel [kd,kh,kw,Cin,Cout]."""
    B, D, H, W, _ = x.shape
    k = kernel.shape[0]
    pd, ph, pw = _same_pad(D, k, stride), _same_pad(H, k, stride), _same_pad(W, k, stride)
    t = TF.pad(_to_t(x), (pw[0], pw[1], ph[0], ph[1], pd[0], pd[1]))
    w = torch.from_numpy(np.ascontiguousarray(kernel, dtype=F32)).permute(4, 3, 0, 1, 2).contiguous()
    return _to_n(TF.conv3d(t, w, stride=stride))


def deconv3d(x, kernel, stride=2):
    """tf.layers.conv3d_transpose, padding SAME, k=3, no bias.  kernel
    [kd,kh,kw,Cout,Cin]; out[s*i + k] += in[i]*w[k], cropped to [0, s*n)."""
    B, D, H, W, _ = x.shape
    w = torch.from_numpy(np.ascontiguousarray(kernel, dtype=F32)).permute(4, 3, 0, 1, 2).contiguous()
    y = TF.conv_transpose3d(_to_t(x), w, stride=stride)
    k = kernel.shape[0]
    # SAME crop: total = k - s (for k >= s), before = total // 2  (k=3,s=2 -> (0,1))
    total = max(k - stride, 0)
    b = total // 2
    y = y[:, :, b:b + D * stride, b:b + H * stride, b:b + W * stride]
    return _to_n(y)


def batch_norm_train(x):
    """tf.layers.batch_normalization(training=True, center=False, scale=False)
    on a 5-D tensor (non-fused path): moments over all axes but the last, biased
    variance, y = x*inv - mean*inv with inv = rsqrt(var + eps)."""
    x = np.asarray(x, dtype=F32)
    axes = tuple(range(x.ndim - 1))
    mean = x.mean(axis=axes, dtype=F32)
    var = np.mean(np.square(x - mean), axis=axes, dtype=F32)
    inv = (F32(1.0) / np.sqrt(var + F32(BN_EPS))).astype(F32)
    return x * inv - mean * inv


def relu(x):
    return np.maximum(x, F32(0))


class Network(object):
    """Minimal stand-in for cnn_wrapper/network.py:Network (feed / layer LUT /
    get_output / get_output_by_name), eager."""

    def __init__(self, inputs, weights, is_training=True, reuse=None):
        self.layers = dict(inputs)
        self.weights = weights
        self.terminals = []
        self.setup()

    def feed(self, *names):
        self.terminals = [self.layers[n] if isinstance(n, str) else n for n in names]
        return self

    def _done(self, name, out):
        self.layers[name] = out
        self.terminals = [out]
        return self

    def get_output(self):
        return self.terminals[-1]

    def get_output_by_name(self, name):
        return self.layers[name]

    # layers -----------------------------------------------------------------
    def conv(self, k, filters, stride, name, relu=True):
        out = conv3d(self.terminals[0], self.weights[name + '/kernel'], stride)
        if relu:
            out = np.maximum(out, F32(0))
        return self._done(name, out)

    def conv_bn(self, k, filters, stride, name):
        out = conv3d(self.terminals[0], self.weights[name + '/conv3d/kernel'], stride)
        return self._done(name, relu(batch_norm_train(out)))

    def deconv_bn(self, k, filters, stride, name):
        out = deconv3d(self.terminals[0], self.weights[name + '/conv3d_transpose/kernel'], stride)
        return self._done(name, relu(batch_norm_train(out)))

    def add(self, name):
        out = self.terminals[0]
        for t in self.terminals[1:]:
            out = out + t
        return self._done(name, out)

    def attention_aggregation(self, name):
        """network.py:379-408 with second_weight=True, relu=True, biased=False.
        input (B,D,H,W,C,N) -> (B,D,H,W,C)."""
        x = self.terminals[0]
        n_view = x.shape[-1]
        w_u = self.weights[name + '/attention_activation/weight_unique']
        w_s = self.weights[name + '/attention_activation/weight_shared']
        shared = [relu(conv3d(x[..., n], w_s)) for n in range(n_view)]       # tf.scan #1
        shared_sum = shared[0]
        for s in shared[1:]:
            shared_sum = shared_sum + s
        act = [(relu(conv3d(x[..., n], w_u)) - shared[n]) + shared_sum for n in range(n_view)]
        act = np.stack(act, axis=-1)
        m = act.max(axis=-1, keepdims=True)
        e = np.exp(act - m)
        score = e / e.sum(axis=-1, keepdims=True, dtype=F32)
        return self._done(name, (score * x).sum(axis=-1, dtype=F32))


class StackedUNet_prob(Network):
    """cnn_wrapper/atvsnet.py:100-192 (and StackedUNet :5-96 = same without the last conv)."""
    with_prob = True

    def setup(self):
        bf = 8
        for b in range(3):
            p = 'conv_b%d' % b
            src = 'data' if b == 0 else p + '_0_0'
            (self.feed(src).conv_bn(3, bf * 2, 2, name=p + '_1_0')
                 .conv_bn(3, bf * 4, 2, name=p + '_2_0')
                 .conv_bn(3, bf * 8, 2, name=p + '_3_0'))
            self.feed(src).conv_bn(3, bf, 1, name=p + '_0_1')
            if b == 0:
                self.feed(p + '_1_0').conv_bn(3, bf * 2, 1, name=p + '_1_1')
                self.feed(p + '_2_0').conv_bn(3, bf * 4, 1, name=p + '_2_1')
            else:
                q = 'conv_b%d' % (b - 1)
                (self.feed(p + '_1_0', q + '_5_0').add(name=p + '_1_1_concat')
                     .conv_bn(3, bf * 2, 1, name=p + '_1_1'))
                (self.feed(p + '_2_0', q + '_4_0').add(name=p + '_2_1_concat')
                     .conv_bn(3, bf * 4, 1, name=p + '_2_1'))
            (self.feed(p + '_3_0').conv_bn(3, bf * 8, 1, name=p + '_3_1')
                 .deconv_bn(3, bf * 4, 2, name=p + '_4_0'))
            extra2 = [] if b == 0 else ['conv_b0_2_1']
            extra1 = [] if b == 0 else ['conv_b0_1_1']
            (self.feed(p + '_4_0', p + '_2_1', *extra2).add(name=p + '_4_1')
                 .deconv_bn(3, bf * 2, 2, name=p + '_5_0'))
            (self.feed(p + '_5_0', p + '_1_1', *extra1).add(name=p + '_5_1')
                 .deconv_bn(3, bf, 2, name=p + '_6_0'))
            nxt = 'conv_b%d_0_0' % (b + 1) if b < 2 else 'conv_b2_6_1'
            self.feed(p + '_6_0', p + '_0_1').add(name=nxt)
        if self.with_prob:
            self.feed('conv_b2_6_1').conv(3, 1, 1, relu=False, name='conv_b2_6_2')


class StackedUNet(StackedUNet_prob):
    with_prob = False


class AttAggregation_keepchannel(Network):
    scope = 'attention_aggregate'

    def setup(self):
        self.feed('data').attention_aggregation(name=self.scope)


class AttAggregation(Network):
    scope, out = 'attention_aggregate', 'attention_prob_vol'

    def setup(self):
        self.feed('data').attention_aggregation(name=self.scope).conv(3, 1, 1, relu=False, name=self.out)


class AttAggregation_refine_keepchannel(AttAggregation_keepchannel):
    scope = 'attention_aggregate_refine'


class AttAggregation_refine(AttAggregation):
    scope, out = 'attention_aggregate_refine', 'attention_prob_vol_refine'


class OutputConv(Network):
    out = 'attention_prob_vol'

    def setup(self):
        self.feed('data').conv(3, 1, 1, relu=False, name=self.out)


class OutputConv_refine(OutputConv):
    out = 'attention_prob_vol_refine'
