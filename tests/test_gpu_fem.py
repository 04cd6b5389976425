"""GPU: the fp32 CUDA path of the 2-D feature extraction module (a-tvsnet_b200/fem.py -> csrc/fem2d.cu through the C ABI)
against the CPU oracle (oracle/fem.py) and the reference-graph golden vectors (tests/golden/reference_golden_fem.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def rel(a, b):
    # layers are O(1) (normalised): an absolute floor of 1 keeps the all-zero SPP branch of a 1x1 pooled map (batch
    # norm of a single element is exactly 0 in the reference) from turning 1e-6 of rounding into an infinite ratio
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1.0))


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope='module')
def A():
    import atvsnet_b200 as A_
    return A_


@pytest.fixture
def fp32_path(A):
    """the fp32 CUDA-core parity path of the FEM (the default precision 'fp16' routes the stride-1 convolutions to the
    tensor cores)."""
    A.FLAGS.precision = 'fp32'
    yield
    A.FLAGS.precision = A.flags.DEFAULT_PRECISION


@pytest.fixture
def tensor_fem(A):
    A.FLAGS.fem_tensor = True
    yield
    A.FLAGS.fem_tensor = False


CONV_CASES = [  # cin, cout, k, stride, rate, mode, shape(H, W)
    (3, 32, 3, 2, 1, 'same', (30, 44)), (32, 32, 3, 1, 1, 'same', (17, 23)), (32, 64, 1, 2, 1, 'same', (16, 24)),
    (64, 64, 3, 2, 1, 'explicit', (16, 24)), (128, 128, 3, 1, 2, 'same', (12, 20)), (128, 128, 3, 1, 4, 'same', (9, 33)),
    (320, 128, 3, 1, 1, 'same', (8, 12)), (128, 32, 1, 1, 1, 'same', (8, 40)), (128, 32, 3, 1, 1, 'same', (2, 3)),
    (20, 24, 3, 1, 1, 'same', (5, 70))]


@pytest.mark.parametrize('cin,cout,k,stride,rate,mode,hw', CONV_CASES)
def test_conv2d_fp32(A, fp32_path, cin, cout, k, stride, rate, mode, hw):
    from oracle import fem as ofem
    rng = np.random.default_rng(cin * 7 + cout)
    x = rng.standard_normal((2,) + hw + (cin,)).astype(np.float32)
    w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    if mode == 'explicit':
        ref = ofem.conv2d(np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0))), w, stride, rate, 'VALID', b)
        got = A.fem.conv2d(cu(x), cu(w), stride, rate, cu(b), relu=False, explicit_pad=(1, 1))
    else:
        ref = ofem.conv2d(x, w, stride, rate, 'SAME', b)
        got = A.fem.conv2d(cu(x), cu(w), stride, rate, cu(b))
    assert tuple(got.shape) == ref.shape
    assert rel(got.cpu().numpy(), ref) < 2e-5
    got_r = A.fem.conv2d(cu(x), cu(w), stride, rate, None, relu=True, explicit_pad=(1, 1) if mode == 'explicit' else None)
    ref_r = np.maximum(ref - b, 0)
    assert rel(got_r.cpu().numpy(), ref_r) < 2e-5


def test_bn_pool_resize(A, fp32_path):
    from oracle import fem as ofem
    rng = np.random.default_rng(5)
    for C in (32, 64, 128, 20):
        x = (rng.standard_normal((1, 13, 21, C)) * 3 + 1).astype(np.float32)
        beta = rng.standard_normal(C).astype(np.float32)
        assert rel(A.fem.batch_norm(cu(x), cu(beta), True).cpu().numpy(), np.maximum(ofem.batch_norm_train(x, beta), 0)) < 2e-5
        assert rel(A.fem.batch_norm(cu(x)).cpu().numpy(), ofem.batch_norm_train(x)) < 2e-5
    x = rng.standard_normal((2, 24, 32, 16)).astype(np.float32)
    for k in (64, 32, 16, 8, 5):
        assert rel(A.fem.avg_pool(cu(x), k, k).cpu().numpy(), ofem.avg_pool_same(x, k, k)) < 1e-5
    for (hi, wi, ho, wo) in ((1, 1, 24, 32), (2, 3, 24, 32), (3, 4, 7, 9), (6, 8, 6, 8)):
        s = rng.standard_normal((2, hi, wi, 8)).astype(np.float32)
        assert rel(A.fem.image_resize(cu(s), ho, wo).cpu().numpy(), ofem.resize_bilinear_align(s, ho, wo)) < 1e-6


TC2D_CASES = [  # cin, cout, k, rate, (H, W)
    (32, 32, 3, 1, (40, 56)), (32, 32, 1, 1, (17, 23)), (32, 64, 1, 1, (16, 24)), (64, 64, 3, 1, (24, 40)),
    (64, 128, 1, 1, (12, 20)), (128, 128, 3, 2, (12, 20)), (128, 128, 3, 4, (9, 33)), (128, 128, 1, 1, (32, 40)),
    (320, 128, 3, 1, (16, 24)), (128, 32, 1, 1, (8, 40)), (64, 64, 3, 1, (128, 160))]


@pytest.mark.parametrize('cin,cout,k,rate,hw', TC2D_CASES)
def test_conv2d_tensor_core(A, tensor_fem, cin, cout, k, rate, hw):
    """the FEM's stride-1 convolutions on tcgen05 (atvs_conv2d_tc: channel-chunked taps, dilation, bias / ReLU epilogue,
    fp32 | fp16 output, moments) against the oracle conv2d on fp16-rounded operands."""
    from oracle import fem as ofem
    rng = np.random.default_rng(cin * 5 + cout + k + rate)
    x = rng.standard_normal((2,) + hw + (cin,)).astype(np.float32)
    w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    xh = torch.from_numpy(x).half()
    wh = torch.from_numpy(w).half().float()
    ref = ofem.conv2d(xh.float().numpy(), wh.numpy(), 1, rate, 'SAME', b)
    A.variables.packed_cache().clear()
    assert A.fem.tc_supported(xh, wh, 1, None)
    got = A.fem.conv2d_tc(xh.cuda(), wh.cuda(), rate, cu(b))
    assert tuple(got.shape) == ref.shape and got.dtype == torch.float32
    assert rel(got.cpu().numpy(), ref) < 1e-4
    stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    got_r = A.fem.conv2d_tc(xh.cuda(), wh.cuda(), rate, cu(b), relu=True, out_dtype=torch.float16, stats=stats)
    ref_r = np.maximum(ref, 0)
    assert got_r.dtype == torch.float16
    assert np.abs(got_r.float().cpu().numpy() - ref_r).max() <= 2.0 ** -10 * np.abs(ref_r).max()
    flat = ref_r.reshape(-1, cout).astype(np.float64)
    st = stats.cpu().numpy()
    assert np.allclose(st[:cout], flat.sum(0), rtol=1e-3, atol=5e-2) and np.allclose(st[cout:], (flat ** 2).sum(0), rtol=1e-3, atol=5e-2)
    nob = A.fem.conv2d_tc(xh.cuda(), wh.cuda(), rate)
    assert rel(nob.cpu().numpy(), ref - b) < 1e-4


def test_resnet_ds2_spp_tensor_path_vs_golden(A, tensor_fem):
    """whole FEM with the stride-1 convolutions on the tensor cores (fp16 operands, fp32 residual stream and batch
    statistics) against the reference-graph golden vectors: ~40 layers of 11-bit operand rounding."""
    from gen_common import fem_weights
    gold = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden_fem.npz')))
    A.variables.load_weights(fem_weights(7))
    assert A.fem.tensor_path()
    out, layers = A.fem.ResNetDS2SPP(cu(gold['image']), return_layers=True)
    torch.cuda.synchronize()
    errs = {nm: rel(layers[nm].cpu().numpy(), gold[nm]) for nm in ('conv0_2', 'conv0_x', 'conv1_x', 'conv2_x', 'conv3_x', 'fusion0')}
    errs['feature'] = rel(out.cpu().numpy(), gold['feature'])
    print("FEM tensor path, max rel err per layer:", {k: "%.2e" % v for k, v in errs.items()})
    assert max(errs.values()) < 1e-2, errs
    mean_rel = float(np.abs(out.cpu().numpy() - gold['feature']).mean() / np.abs(gold['feature']).mean())
    print("FEM tensor path: mean |err| / mean |feature| = %.2e" % mean_rel)
    assert mean_rel < 1.5e-2


def test_resnet_ds2_spp_matches_reference_graph_golden(A, fp32_path):
    from gen_common import fem_weights
    gold = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden_fem.npz')))
    A.variables.load_weights(fem_weights(7))
    out, layers = A.fem.ResNetDS2SPP(cu(gold['image']), return_layers=True)
    torch.cuda.synchronize()
    for nm in ('conv0_2', 'conv0_x', 'conv1_x', 'conv2_x', 'conv3_x', 'branch_0', 'branch_3', 'fusion0'):
        assert rel(layers[nm].cpu().numpy(), gold[nm]) < 1e-4, (nm, rel(layers[nm].cpu().numpy(), gold[nm]))
    assert tuple(out.shape) == (1, 24, 32, 32)
    assert rel(out.cpu().numpy(), gold['feature']) < 1e-4


def test_images_to_depth_end_to_end(A):
    """images -> FEM -> stage I + II: fp32 CUDA path against the CPU oracle on the same weights."""
    from gen_common import fem_weights
    from oracle import fem as ofem
    from oracle import model as om
    rng = np.random.default_rng(9)
    nv, H, W, D = 3, 64, 96, 16
    imgs = (127.5 + 50 * rng.standard_normal((1, nv, H, W, 3))).clip(0, 255).astype(np.float32)
    wf = fem_weights(7)
    wn = A.variables.synthetic_weights(seed=11, logit_gain=2.0)
    cams = A.synthetic.orbit_cams(nv, H // 4, W // 4, D)[None]
    feats_ref = np.stack([ofem.ResNetDS2SPP(imgs[:, n], wf)[0] for n in range(nv)])[None]
    ref = om.run_multiview_stage12(feats_ref, cams, D, wn, siamese=False)
    allw = dict(wn)
    allw.update(wf)
    A.variables.load_weights(allw)
    A.FLAGS.precision = 'fp32'
    try:
        feats = A.fem.extract_features(cu(imgs))
        assert rel(feats.cpu().numpy(), feats_ref) < 2e-4
        out = A.pipeline.run_multiview(feats, cu(cams), D, siamese=False)
    finally:
        A.FLAGS.precision = A.flags.DEFAULT_PRECISION
    rng_ = float((D - 1) * cams[0, 0, 1, 3, 1])
    assert float(np.abs(out['depth_up'].cpu().numpy() - ref['depth_agg_init_up']).mean()) / rng_ < 1e-3


def test_extract_features_batched_views_equal_per_view_towers(A):
    """extract_features runs the towers of all views as ONE batch (view-major) with batch statistics per view
    (fem._GROUPS): the same numbers as one tower per (B,H,W,3) view slice, the form of model.py:420-425 - for B = 1
    (example.py) and for B = 2 (statistics over the two samples of a view, never across views)."""
    from gen_common import fem_weights
    A.variables.load_weights(fem_weights(5))
    rng = np.random.default_rng(21)
    A.FLAGS.precision = 'fp32'
    try:
        for B, nv, H, W in ((1, 4, 64, 96), (2, 3, 48, 64)):
            imgs = cu((127.5 + 50 * rng.standard_normal((B, nv, H, W, 3))).clip(0, 255).astype(np.float32))
            got = A.fem.extract_features(imgs)
            A.fem.BATCH_VIEWS = False
            try:
                ref = A.fem.extract_features(imgs)
            finally:
                A.fem.BATCH_VIEWS = True
            assert tuple(got.shape) == (B, nv, H // 4, W // 4, 32)
            assert rel(got.cpu().numpy(), ref.cpu().numpy()) < 1e-6
    finally:
        A.FLAGS.precision = A.flags.DEFAULT_PRECISION


def test_channel_moments_few_channels(A):
    """atvs_channel_moments on channel counts that do not divide 256 (the 3-channel image in front of the shallow feature
    net's first BN takes the pixel-walking kernel): sums against NumPy fp64."""
    rng = np.random.default_rng(3)
    for C, count in ((3, 70001), (5, 4099), (6, 257), (7, 1), (3, 327680)):
        x = rng.standard_normal((count, C)).astype(np.float32) * 3 + 1
        st = torch.zeros(2 * C, dtype=torch.float64, device='cuda')
        A._lib.call("atvs_channel_moments", A._lib.ptr(cu(x)), count, C, A._lib.ptr(st), A._lib.stream())
        got = st.cpu().numpy()
        x64 = x.astype(np.float64)
        assert np.allclose(got[:C], x64.sum(0), rtol=1e-12, atol=1e-9)
        assert np.allclose(got[C:], (x64 ** 2).sum(0), rtol=1e-12, atol=1e-9)
