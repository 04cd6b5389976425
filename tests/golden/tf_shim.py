"""A NumPy stand-in for the handful of TensorFlow-1.5 primitives that the reference's
hot-path source files call, so that those files can be executed UNMODIFIED under
Python 3 in the build container (make_golden.py).  TEST INFRASTRUCTURE ONLY; it is used
once, offline, to generate tests/golden/*.npz and never on the GPU box.

Everything is eager: a "tensor" is a float32/int32/bool ``numpy.ndarray`` (subclass ``T``
that adds ``get_shape()``).  Leaf semantics restated from the TF 1.5 documentation:

* ``matmul``           fp32 products accumulated over k in index order (no BLAS/FMA)
* ``gather_nd``        integer fancy indexing
* ``layers.conv3d``    NDHWC, kernel [kd,kh,kw,Cin,Cout], 'SAME' = pad_total
                       max((ceil(n/s)-1)*s+k-n,0), floor(total/2) in front; plain sum over taps
* ``layers.conv3d_transpose``  gradient of conv3d wrt its input: out[s*i+k-pad] += in[i]*w[k]
                       with the SAME padding of the forward conv, output length s*n,
                       kernel [kd,kh,kw,Cout,Cin]
* ``layers.batch_normalization(training=True)``  moments over all but the last axis (biased),
                       ``x*inv - mean*inv`` with ``inv = rsqrt(var+1e-3)`` (+beta if center)
* ``nn.softmax``       exp(x-max)/sum
* ``image.resize_images(BILINEAR, align_corners=True)``  src = dst*(in-1)/(out-1), lerp
* ``scan``             python loop over the leading axis
* variables            looked up by ``variable_scope``-qualified name in ``VARIABLES``
"""
import sys
import types

import numpy as np

F32 = np.float32
VARIABLES = {}          # name -> np.ndarray (TF layouts)
_SCOPE = []


class T(np.ndarray):
    def get_shape(self):
        return _Shape(self.shape)


class _Dim(int):
    @property
    def value(self):
        return int(self)


class _Shape(list):
    def __init__(self, shp):
        list.__init__(self, [_Dim(s) for s in shp])

    def as_list(self):
        return [int(s) for s in self]


def _t(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    if a.dtype == np.float64:
        a = a.astype(F32)
    if a.dtype == np.int64:
        a = a.astype(np.int32)
    return a.view(T)


def _dt(d):
    if d in ('float32', 'float', F32, np.dtype('float32')):
        return F32
    if d in ('int32', np.int32, np.dtype('int32')):
        return np.int32
    if d in ('bool', np.bool_, np.dtype('bool')):
        return np.bool_
    return np.dtype(d).type


class _Ctx(object):
    def __init__(self, name=None, push=False):
        self.name, self.push = name, push

    def __enter__(self):
        if self.push:
            _SCOPE.append(self.name)
        return self.name

    def __exit__(self, *a):
        if self.push:
            _SCOPE.pop()
        return False


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2, out


def _conv3d(x, w, stride):
    x, w = np.asarray(x, F32), np.asarray(w, F32)
    B, D, H, W, Ci = x.shape
    k = w.shape[0]
    (pd0, pd1, Do), (ph0, ph1, Ho), (pw0, pw1, Wo) = (_same_pad(n, k, stride) for n in (D, H, W))
    xp = np.pad(x, ((0, 0), (pd0, pd1), (ph0, ph1), (pw0, pw1), (0, 0)))
    out = np.zeros((B, Do, Ho, Wo, w.shape[-1]), F32)
    for a in range(k):
        for b in range(k):
            for c in range(k):
                sl = xp[:, a:a + (Do - 1) * stride + 1:stride, b:b + (Ho - 1) * stride + 1:stride,
                        c:c + (Wo - 1) * stride + 1:stride, :]
                out += np.einsum('bdhwi,io->bdhwo', sl, w[a, b, c], optimize=False).astype(F32)
    return _t(out)


def _conv3d_transpose(x, w, stride):
    """w [kd,kh,kw,Cout,Cin]."""
    x, w = np.asarray(x, F32), np.asarray(w, F32)
    B, D, H, W, Ci = x.shape
    k = w.shape[0]
    Co = w.shape[3]
    full = np.zeros((B, (D - 1) * stride + k, (H - 1) * stride + k, (W - 1) * stride + k, Co), F32)
    for a in range(k):
        for b in range(k):
            for c in range(k):
                full[:, a:a + (D - 1) * stride + 1:stride, b:b + (H - 1) * stride + 1:stride,
                     c:c + (W - 1) * stride + 1:stride, :] += \
                    np.einsum('bdhwi,oi->bdhwo', x, w[a, b, c], optimize=False).astype(F32)
    # forward conv over an input of length s*n has SAME pad (p0, p1); its gradient crops p0 in front
    pads = [_same_pad(n * stride, k, stride)[0] for n in (D, H, W)]
    return _t(full[:, pads[0]:pads[0] + D * stride, pads[1]:pads[1] + H * stride, pads[2]:pads[2] + W * stride, :])


def _var(name, shape=None):
    full = '/'.join(_SCOPE + [name])
    if full not in VARIABLES:
        raise KeyError('tf_shim: variable %r not provided' % full)
    v = np.asarray(VARIABLES[full], F32)
    if shape is not None and tuple(int(s) for s in shape) != v.shape:
        raise ValueError('tf_shim: variable %r has shape %s, graph wants %s' % (full, v.shape, tuple(shape)))
    return _t(v)


def _layers_conv3d(inputs, filters, kernel_size, strides=1, activation=None, use_bias=False, padding='SAME',
                   trainable=True, reuse=None, name=None, kernel_initializer=None, dilation_rate=1):
    assert padding == 'SAME' and not use_bias and dilation_rate == 1
    with _Ctx(name or 'conv3d', push=True):
        w = _var2('kernel', (kernel_size,) * 3 + (inputs.shape[-1], filters), 'kernel')
    y = _conv3d(inputs, w, strides)
    return activation(y) if activation else y


def _layers_conv3d_transpose(inputs, filters, kernel_size, strides=1, activation=None, use_bias=False,
                             padding='SAME', trainable=True, reuse=None, name=None, kernel_initializer=None):
    assert padding == 'SAME' and not use_bias
    with _Ctx(name or 'conv3d_transpose', push=True):
        w = _var2('kernel', (kernel_size,) * 3 + (filters, inputs.shape[-1]), 'kernel')
    y = _conv3d_transpose(inputs, w, strides)
    return activation(y) if activation else y


def _batch_norm(x, center=True, scale=True, training=False, fused=None, trainable=True, reuse=None, name=None,
                epsilon=1e-3, axis=-1):
    assert training, 'the reference builds every net with is_training=True'
    x = np.asarray(x, F32)
    axes = tuple(range(x.ndim - 1))
    mean = x.mean(axis=axes, dtype=F32)
    var = np.mean(np.square(x - mean), axis=axes, dtype=F32)
    inv = (F32(1) / np.sqrt(var + F32(epsilon))).astype(F32)
    y = x * inv - mean * inv
    if center:
        with _Ctx(name or 'batch_normalization', push=True):
            y = y + _var('beta')
    return _t(y)


def _softmax(x, axis=-1, name=None, dim=None):
    if dim is not None:
        axis = dim
    x = np.asarray(x, F32)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return _t(e / e.sum(axis=axis, keepdims=True, dtype=F32))


def _matmul(a, b):
    a, b = np.asarray(a), np.asarray(b)
    K = a.shape[-1]
    out = None
    for k in range(K):
        term = a[..., :, k:k + 1] * b[..., k:k + 1, :]
        out = term if out is None else out + term
    return _t(out)


def _inv(m):
    # tf.matrix_inverse: LU with partial pivoting in fp32 (LAPACK sgetrf/sgetri here)
    return _t(np.linalg.inv(np.asarray(m, F32)))


def _slice(x, begin, size):
    x = np.asarray(x)
    idx = tuple(slice(int(b), None if int(s) == -1 else int(b) + int(s)) for b, s in zip(begin, size))
    return _t(x[idx])


def _gather_nd(params, indices):
    params, indices = np.asarray(params), np.asarray(indices)
    return _t(params[tuple(indices[..., i] for i in range(indices.shape[-1]))])


def _linspace(start, stop, num):
    num = int(num)
    start, stop = F32(start), F32(stop)
    if num == 1:
        return _t(np.array([start], F32))
    step = (stop - start) / F32(num - 1)
    return _t(start + np.arange(num, dtype=F32) * step)


def _resize_images(images, size, method=None, align_corners=False):
    assert align_corners
    v = np.asarray(images, F32)
    B, H, W, C = v.shape
    Ho, Wo = int(size[0]), int(size[1])

    def axis(n_in, n_out):
        scale = F32(n_in - 1) / F32(n_out - 1) if n_out > 1 else F32(0)
        src = np.arange(n_out, dtype=F32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        return lo, hi, (src - lo.astype(F32)).astype(F32)

    y0, y1, fy = axis(H, Ho)
    x0, x1, fx = axis(W, Wo)
    fy = fy[None, :, None, None]
    fx = fx[None, None, :, None]
    tl, tr = v[:, y0][:, :, x0], v[:, y0][:, :, x1]
    bl, br = v[:, y1][:, :, x0], v[:, y1][:, :, x1]
    top = tl + (tr - tl) * fx
    bot = bl + (br - bl) * fx
    return _t(top + (bot - top) * fy)


def _scan(fn, elems, initializer=None):
    n = elems[0].shape[0]
    outs, prev = [], initializer
    for i in range(n):
        prev = fn(prev, tuple(_t(e[i]) for e in elems))
        outs.append(prev)
    return _t(np.stack(outs, axis=0))


def _cast(x, dtype=None, **kw):
    dtype = _dt(dtype if dtype is not None else kw.get('dtype'))
    a = np.asarray(x)
    if np.issubdtype(dtype, np.integer) and np.issubdtype(a.dtype, np.floating):
        with np.errstate(invalid='ignore'):
            a = np.where(np.isfinite(a), a, 0)
    return _t(a.astype(dtype))


def _eye(n, batch_shape=None):
    e = np.eye(n, dtype=F32)
    if batch_shape is not None:
        e = np.broadcast_to(e, tuple(int(b) for b in batch_shape) + (n, n)).copy()
    return _t(e)


def _reduce(fn):
    def f(x, axis=None, keepdims=False, name=None, keep_dims=None):
        if keep_dims is not None:
            keepdims = keep_dims
        return _t(fn(np.asarray(x), axis=axis, keepdims=keepdims))
    return f


def install(flags=None):
    """Create the fake ``tensorflow`` module tree and register it in sys.modules."""
    tf = types.ModuleType('tensorflow')
    FLAGS = types.SimpleNamespace(inverse_depth=True, batch_size=1, view_num=5, max_d=128)
    if flags:
        FLAGS.__dict__.update(flags)
    tf.app = types.SimpleNamespace(flags=types.SimpleNamespace(FLAGS=FLAGS))
    tf.float32, tf.int32, tf.bool = F32, np.int32, np.bool_
    tf.AUTO_REUSE = 'AUTO_REUSE'
    tf.name_scope = lambda name=None, *a, **k: _Ctx(name, push=False)
    tf.variable_scope = lambda name=None, *a, **k: _Ctx(name, push=True)
    tf.get_variable = lambda name, shape=None, initializer=None, trainable=True: _var(name, shape)
    tf.zeros_initializer = lambda *a, **k: None
    tf.constant = lambda v, dtype=None, **k: _t(v, _dt(dtype) if dtype is not None else None)
    tf.shape = lambda x: _t(np.array(np.asarray(x).shape, np.int32))
    tf.cast = _cast
    tf.linspace = _linspace
    tf.meshgrid = lambda *xs: [_t(m) for m in np.meshgrid(*[np.asarray(x) for x in xs])]
    tf.reshape = lambda x, shape, name=None: _t(np.reshape(np.asarray(x), [int(v) for v in np.asarray(shape).reshape(-1)]))
    tf.ones_like = lambda x: _t(np.ones_like(np.asarray(x)))
    tf.ones = lambda shape, dtype='float32': _t(np.ones([int(s) for s in shape], _dt(dtype)))
    tf.zeros = lambda shape, dtype='float32': _t(np.zeros([int(s) for s in shape], _dt(dtype)))
    tf.concat = lambda values, axis, name=None: _t(np.concatenate([np.asarray(v) for v in values], axis=axis))
    tf.stack = lambda values, axis=0, name=None: _t(np.stack([np.asarray(v) for v in values], axis=axis))
    tf.unstack = lambda x, axis=0: [_t(np.take(np.asarray(x), i, axis=axis)) for i in range(np.asarray(x).shape[axis])]
    tf.matmul = _matmul
    tf.matrix_inverse = _inv
    tf.transpose = lambda x, perm=None, name=None, conjugate=False: _t(np.transpose(np.asarray(x), perm))
    tf.tile = lambda x, multiples, name=None: _t(np.tile(np.asarray(x), [int(m) for m in multiples]))
    tf.expand_dims = lambda x, axis, name=None: _t(np.expand_dims(np.asarray(x), axis))
    tf.squeeze = lambda x, axis=None, name=None, squeeze_dims=None: _t(
        np.squeeze(np.asarray(x), axis=axis if axis is not None else squeeze_dims))
    tf.slice = _slice
    tf.range = lambda *a, **k: _t(np.arange(*[int(v) for v in a], dtype=np.int32))
    tf.gather_nd = _gather_nd
    tf.logical_and = lambda a, b: _t(np.logical_and(a, b))
    tf.logical_not = lambda a: _t(np.logical_not(a))
    tf.greater_equal = lambda a, b: _t(np.greater_equal(a, b))
    tf.greater = lambda a, b: _t(np.greater(a, b))
    tf.less = lambda a, b: _t(np.less(a, b))
    tf.equal = lambda a, b: _t(np.equal(a, b))
    tf.is_nan = lambda a: _t(np.isnan(a))
    tf.round = lambda a: _t(np.rint(a))
    tf.floor = lambda a: _t(np.floor(a))
    tf.ceil = lambda a: _t(np.ceil(a))
    tf.abs = lambda a: _t(np.abs(a))
    tf.multiply = lambda a, b, name=None: _t(np.asarray(a) * np.asarray(b))
    tf.scalar_mul = lambda s, a: _t(F32(s) * np.asarray(a))
    tf.add = lambda a, b, name=None: _t(np.asarray(a) + np.asarray(b))
    tf.subtract = lambda a, b, name=None: _t(np.asarray(a) - np.asarray(b))
    tf.div = lambda a, b, name=None: _t(np.asarray(a) / np.asarray(b))
    tf.divide = tf.div
    tf.reciprocal = lambda a: _t(F32(1) / np.asarray(a))

    def add_n(xs, name=None):
        out = np.asarray(xs[0])
        for x in xs[1:]:
            out = out + np.asarray(x)
        return _t(out)
    tf.add_n = add_n
    tf.clip_by_value = lambda x, lo, hi: _t(np.clip(np.asarray(x), lo, hi))
    tf.eye = _eye
    tf.reduce_sum = _reduce(np.sum)
    tf.reduce_max = _reduce(np.max)
    tf.reduce_mean = _reduce(np.mean)
    tf.scan = _scan
    tf.nn = types.SimpleNamespace(
        relu=lambda x, name=None: _t(np.maximum(np.asarray(x), F32(0))),
        softmax=_softmax,
        conv3d=lambda x, k, strides, padding: _conv3d(x, k, int(strides[1])),
        bias_add=lambda x, b: _t(np.asarray(x) + np.asarray(b)))
    tf.layers = types.SimpleNamespace(conv3d=_layers_conv3d, conv3d_transpose=_layers_conv3d_transpose,
                                      batch_normalization=_batch_norm)
    tf.image = types.SimpleNamespace(resize_images=_resize_images,
                                     ResizeMethod=types.SimpleNamespace(BILINEAR='bilinear',
                                                                        NEAREST_NEIGHBOR='nearest'))
    contrib = types.ModuleType('tensorflow.contrib')
    contrib.slim = types.SimpleNamespace()
    contrib.layers = types.SimpleNamespace(xavier_initializer=lambda *a, **k: None)
    tf.contrib = contrib
    sys.modules['tensorflow'] = tf
    sys.modules['tensorflow.contrib'] = contrib
    return tf


# ------------------------------------------------------------------ 2-D primitives (FEM ResNetDS2SPP: cnn_wrapper/atvsnet.py:254-292,
# network.py:552-616, 650-671).  NumPy einsum / slicing, independent of the torch-CPU oracle.
AUTO_VARS = None      # when a dict: variables the graph asks for and that were not provided are CREATED (seeded) and
                      # recorded as name -> shape, so that the variable list itself comes from the reference's code


def _var2(name, shape, kind):
    full = '/'.join(_SCOPE + [name])
    if full not in VARIABLES and AUTO_VARS is not None:
        rng = np.random.default_rng(__import__("zlib").crc32(full.encode()))
        shape = tuple(int(v) for v in shape)
        if kind == 'kernel':
            fan_in = int(np.prod(shape[:-1]))
            VARIABLES[full] = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(F32)
        else:
            VARIABLES[full] = (rng.standard_normal(shape) * 0.1).astype(F32)
        AUTO_VARS[full] = shape
    return _var(name, shape)


def _conv2d(x, w, stride, rate, pads):
    """x (B,H,W,Ci), w [kh,kw,Ci,Co], explicit pads ((top,bottom),(left,right)), dilation `rate`."""
    x, w = np.asarray(x, F32), np.asarray(w, F32)
    B, H, W, Ci = x.shape
    kh, kw = w.shape[0], w.shape[1]
    xp = np.pad(x, ((0, 0), pads[0], pads[1], (0, 0)))
    Hp, Wp = xp.shape[1], xp.shape[2]
    Ho = (Hp - (kh - 1) * rate - 1) // stride + 1
    Wo = (Wp - (kw - 1) * rate - 1) // stride + 1
    out = np.zeros((B, Ho, Wo, w.shape[-1]), F32)
    for a in range(kh):
        for b in range(kw):
            sl = xp[:, a * rate:a * rate + (Ho - 1) * stride + 1:stride, b * rate:b * rate + (Wo - 1) * stride + 1:stride, :]
            out += np.einsum('bhwi,io->bhwo', sl, w[a, b], optimize=False).astype(F32)
    return _t(out)


def _same_pads2d(H, W, kh, kw, stride, rate):
    keh, kew = (kh - 1) * rate + 1, (kw - 1) * rate + 1
    ph, pw = _same_pad(H, keh, stride), _same_pad(W, kew, stride)
    return (ph[0], ph[1]), (pw[0], pw[1])


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (list, tuple)) else (int(v), int(v))


def _layers_conv2d(inputs, filters, kernel_size, strides=1, activation=None, use_bias=False, padding='SAME',
                   trainable=True, reuse=None, name=None, kernel_initializer=None, dilation_rate=1):
    kh, kw = _pair(kernel_size)
    s = _pair(strides)[0]
    r = _pair(dilation_rate)[0]
    with _Ctx(name or 'conv2d', push=True):
        w = _var2('kernel', (kh, kw, inputs.shape[-1], filters), 'kernel')
        b = _var2('bias', (filters,), 'bias') if use_bias else None
    pads = _same_pads2d(inputs.shape[1], inputs.shape[2], kh, kw, s, r) if padding == 'SAME' else ((0, 0), (0, 0))
    y = _conv2d(inputs, w, s, r, pads)
    if b is not None:
        y = _t(np.asarray(y) + np.asarray(b))
    return activation(y) if activation else y


def _slim_conv2d(inputs, num_outputs, kernel_size, stride=1, padding='SAME', rate=1, activation_fn='relu',
                 normalizer_fn=None, reuse=None, trainable=True, weights_initializer=None, scope=None, **kw):
    assert normalizer_fn is None
    kh, kw_ = _pair(kernel_size)
    with _Ctx(scope or 'Conv', push=True):
        w = _var2('weights', (kh, kw_, inputs.shape[-1], num_outputs), 'kernel')
        b = _var2('biases', (num_outputs,), 'bias')
    pads = _same_pads2d(inputs.shape[1], inputs.shape[2], kh, kw_, stride, rate) if padding == 'SAME' else ((0, 0), (0, 0))
    y = np.asarray(_conv2d(inputs, w, stride, rate, pads)) + np.asarray(b)
    if activation_fn == 'relu':                       # slim.conv2d's default activation_fn is tf.nn.relu
        y = np.maximum(y, F32(0))
    elif activation_fn is not None:
        y = np.asarray(activation_fn(_t(y)))
    return _t(y.astype(F32))


def _slim_batch_norm(inputs, activation_fn=None, scope=None, reuse=None, trainable=True, decay=0.999, center=True,
                     scale=False, epsilon=0.001, is_training=True, **kw):
    """slim.batch_norm defaults: center=True, scale=False, epsilon=1e-3, is_training=True (batch statistics)."""
    assert is_training and center and not scale
    x = np.asarray(inputs, F32)
    axes = tuple(range(x.ndim - 1))
    mean = x.mean(axis=axes, dtype=F32)
    var = np.mean(np.square(x - mean), axis=axes, dtype=F32)
    inv = (F32(1) / np.sqrt(var + F32(epsilon))).astype(F32)
    with _Ctx(scope or 'BatchNorm', push=True):
        beta = np.asarray(_var2('beta', (x.shape[-1],), 'bias'))
    y = x * inv - mean * inv + beta
    if activation_fn is not None:
        y = np.asarray(activation_fn(_t(y)))
    return _t(y.astype(F32))


def _slim_max_pool2d(inputs, kernel_size, stride=2, padding='VALID', scope=None):
    kh, kw = _pair(kernel_size)
    assert (kh, kw) == (1, 1) and padding == 'VALID'          # the only use: shortcut subsampling (network.py:576)
    return _t(np.asarray(inputs)[:, ::stride, ::stride, :])


def _avg_pool2d_same(inputs, pool_size, strides, padding='SAME', name=None):
    """tf.layers.average_pooling2d, SAME: windows are clipped to the image and averaged over the VALID elements."""
    assert padding == 'SAME'
    x = np.asarray(inputs, F32)
    B, H, W, C = x.shape
    k, s = _pair(pool_size)[0], _pair(strides)[0]
    (pt, pb, Ho), (pl, pr, Wo) = _same_pad(H, k, s), _same_pad(W, k, s)
    out = np.zeros((B, Ho, Wo, C), F32)
    for i in range(Ho):
        y0, y1 = max(i * s - pt, 0), min(i * s - pt + k, H)
        for j in range(Wo):
            x0, x1 = max(j * s - pl, 0), min(j * s - pl + k, W)
            out[:, i, j, :] = x[:, y0:y1, x0:x1, :].sum(axis=(1, 2), dtype=F32) / F32((y1 - y0) * (x1 - x0))
    return _t(out)


def install_2d(tf):
    """add the 2-D layer set to a module tree made by install()."""
    tf.pad = lambda x, paddings, **k: _t(np.pad(np.asarray(x), [(int(a), int(b)) for a, b in paddings]))
    tf.layers.conv2d = _layers_conv2d
    tf.layers.average_pooling2d = _avg_pool2d_same
    relu = tf.nn.relu
    slim = tf.contrib.slim

    def conv2d(*a, **k):
        if 'activation_fn' not in k:
            k['activation_fn'] = 'relu'
        elif k['activation_fn'] is relu:
            k['activation_fn'] = 'relu'
        return _slim_conv2d(*a, **k)
    slim.conv2d = conv2d
    slim.batch_norm = _slim_batch_norm
    slim.max_pool2d = _slim_max_pool2d
    slim.utils = types.SimpleNamespace(last_dimension=lambda shape, min_rank=1: int(shape[-1]))
    return tf
