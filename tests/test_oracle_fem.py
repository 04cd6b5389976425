"""CPU: the FEM oracle (oracle/fem.py, SURVEY.md 8(f) row N1 - the checker for the next row to be built) against the
reference's own ResNetDS2SPP graph code executed on the TF stand-in (tests/golden/make_golden_fem.py), plus
known-answer checks of the 2-D leaf semantics the two share an assumption about."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from gen_common import fem_variable_shapes, fem_weights  # noqa: E402
from oracle import fem  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden_fem.npz')))


def test_variable_list_matches_checkpoint_convention():
    shapes = fem_variable_shapes()
    assert len(shapes) == 132 and sum(int(np.prod(s)) for s in shapes.values()) == 2020672
    assert shapes['conv0_0/conv2d/kernel'] == (3, 3, 3, 32)
    assert shapes['conv1_x_0/shortcut/weights'] == (1, 1, 32, 64) and 'conv1_x_1/shortcut/weights' not in shapes
    assert shapes['conv1_x/conv2/weights'] == (3, 3, 64, 64)          # last block of a group takes the bare name
    assert 'conv1_x_7/conv1/weights' not in shapes and 'conv1_x_6/conv1/weights' in shapes
    assert shapes['conv2_x_0/preact/beta'] == (64,) and shapes['conv3_x/conv3/biases'] == (128,)
    assert shapes['branch_3_conv/conv2d/kernel'] == (3, 3, 128, 32)
    assert shapes['fusion0/conv2d/kernel'] == (3, 3, 320, 128) and shapes['fusion1/kernel'] == (1, 1, 128, 32)
    assert not any(k.endswith('/bias') for k in shapes)              # conv_bn / conv layers are bias free


def test_fem_oracle_matches_reference_graph(gold):
    w = fem_weights(7)
    out, layers = fem.ResNetDS2SPP(gold['image'], w, return_layers=True)
    assert out.shape == gold['feature'].shape == (1, 24, 32, 32)
    for nm in ('conv0_2', 'conv0_x', 'conv1_x', 'conv2_x', 'conv3_x', 'branch_0', 'branch_3', 'fusion0'):
        assert layers[nm].shape == gold[nm].shape, nm
        assert rel(layers[nm], gold[nm]) < 2e-5, (nm, rel(layers[nm], gold[nm]))
    assert rel(out, gold['feature']) < 2e-5


def test_conv2d_same_padding_hand_cases():
    # stride 2, k 3, even extent: TF SAME pads (0,1): out[i] = x[2i] + x[2i+1] + x[2i+2] with x[n] = 0
    x = np.arange(1, 7, dtype=np.float32).reshape(1, 1, 6, 1)
    k = np.ones((1, 3, 1, 1), np.float32)
    assert fem.conv2d(x, k, stride=2)[0, 0, :, 0].tolist() == [6.0, 12.0, 11.0]
    # the bottleneck's explicit (1,1) pad + VALID differs: out[i] = x[2i-1] + x[2i] + x[2i+1]
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (0, 0)))
    assert fem.conv2d(xp, k, stride=2, padding='VALID')[0, 0, :, 0].tolist() == [3.0, 9.0, 15.0]
    # dilation 2: taps at -2, 0, +2
    assert fem.conv2d(x, k, rate=2)[0, 0, :, 0].tolist() == [4.0, 6.0, 9.0, 12.0, 8.0, 10.0]
    # 1x1 stride 2 = subsampling
    assert fem.conv2d(x, np.ones((1, 1, 1, 1), np.float32), stride=2)[0, 0, :, 0].tolist() == [1.0, 3.0, 5.0]


def test_avg_pool_same_counts_valid_elements_only():
    x = np.ones((1, 5, 7, 2), np.float32)
    y = fem.avg_pool_same(x, 4, 4)
    assert y.shape == (1, 2, 2, 2) and np.allclose(y, 1.0)            # a zero-padded mean would be < 1 at the borders
    x = np.arange(35, dtype=np.float32).reshape(1, 5, 7, 1)
    y = fem.avg_pool_same(x, 64, 64)                                  # window larger than the map: the global mean
    assert y.shape == (1, 1, 1, 1) and np.isclose(y[0, 0, 0, 0], 17.0)


def test_resize_align_corners_and_bn_beta():
    x = np.array([[0.0, 3.0], [6.0, 9.0]], np.float32).reshape(1, 2, 2, 1)
    y = fem.resize_bilinear_align(x, 4, 4)[0, :, :, 0]
    assert np.allclose(y[0], [0, 1, 2, 3]) and np.allclose(y[:, 0], [0, 2, 4, 6]) and np.isclose(y[3, 3], 9.0)
    assert np.allclose(fem.resize_bilinear_align(x[:, :1, :1], 3, 5), 0.0)   # 1x1 source: constant
    rng = np.random.default_rng(0)
    z = rng.standard_normal((2, 6, 5, 3)).astype(np.float32) * 4 + 2
    beta = np.float32([0.5, -1.0, 0.0])
    out = fem.batch_norm_train(z, beta)
    assert np.allclose(out.mean(axis=(0, 1, 2)), beta, atol=1e-5)
    assert np.allclose(out.var(axis=(0, 1, 2)), z.var(axis=(0, 1, 2)) / (z.var(axis=(0, 1, 2)) + 1e-3), rtol=1e-4)


def test_bottleneck_shortcut_kinds():
    w = fem_weights(7)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 8, 12, 32)).astype(np.float32)
    same = fem.bottleneck(x, w, 'conv0_x_0', 32)                      # identity shortcut
    w0 = {k: (np.zeros_like(v) if 'conv3' in k else v) for k, v in w.items()}
    assert np.array_equal(fem.bottleneck(x, w0, 'conv0_x_0', 32), x)
    assert same.shape == x.shape and not np.array_equal(same, x)
    down = fem.bottleneck(x, w, 'conv1_x_0', 64, stride=2)            # projection shortcut, stride 2
    assert down.shape == (1, 4, 6, 64)
    sub = fem.bottleneck(x.repeat(2, axis=-1), w0, 'conv1_x_1', 64, stride=2)     # same depth, stride 2: 1x1 max-pool
    assert np.array_equal(sub, x.repeat(2, axis=-1)[:, ::2, ::2])
