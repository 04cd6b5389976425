"""CPU: known-answer and self-consistency checks of the oracle, using the
bundled example cameras stored in the golden fixture and analytic cases."""
import numpy as np

from oracle import homography_warping as ohw
from oracle import model as om
from oracle import network as onet


def test_identity_pose_is_identity_warp(golden):
    cams = golden['warp_cams']
    H = ohw.get_homographies(cams[0:1], cams[0:1], 8, golden['ds'], golden['di'])
    assert np.allclose(H, np.eye(3)[None, None], atol=2e-5)
    img = golden['warp_img']
    out, mask = ohw.homography_warping(img, np.eye(3, dtype=np.float32)[None], output_mask=True)
    assert np.array_equal(out[:, :-1, :-1], img[:, :-1, :-1])
    assert not out[:, -1].any() and not out[:, :, -1].any()          # strict x < W-1, y < H-1
    assert mask[:, :-1, :-1].all() and not mask[:, -1].any()


def test_plane_sweep_equals_by_depth_at_constant_inverse_depth(golden):
    cams, img = golden['warp_cams'], golden['warp_img']
    ds, di = golden['ds'], golden['di'] * 16
    H = ohw.get_homographies(cams[0:1], cams[1:2], 8, ds, di)
    for d in (0, 5):
        inv = np.full((1,) + img.shape[1:3] + (1,), ds[0] + np.float32(d) * di[0], np.float32)
        a, ma = ohw.homography_warping(img, H[:, d], output_mask=True)
        b, mb = ohw.homography_warping_by_depth(img, cams[0:1], cams[1:2], inv, output_mask=True)
        assert (ma == mb).mean() > 0.995
        ok = (ma & mb)[..., 0]
        assert np.abs(a[ok] - b[ok]).max() < 1e-3 * np.abs(img).max()


def test_pure_translation_analytic():
    """R=I for both cameras, baseline along x: plane d shifts the image by f*b*delta pixels."""
    K = np.float32([[100, 0, 32], [0, 100, 24], [0, 0, 1]])

    def cam(tx):
        c = np.zeros((1, 2, 4, 4), np.float32)
        c[0, 0] = np.eye(4)
        c[0, 0, 0, 3] = tx
        c[0, 1, :3, :3] = K
        return c
    left, right = cam(0.0), cam(-0.5)          # right camera centre at x=+0.5
    H = ohw.get_homographies(left, right, 4, np.float32([0.02]), np.float32([0.02]))
    for d in range(4):
        delta = 0.02 + 0.02 * d
        expect = np.eye(3)
        expect[0, 2] = -100 * 0.5 * delta      # u' = u - f*b*inverse_depth
        assert np.allclose(H[0, d], expect, atol=1e-4)
    x = np.arange(48 * 64, dtype=np.float32).reshape(1, 48, 64, 1)
    out = ohw.homography_warping(x, H[:, 1])   # shift by -2 px exactly
    assert np.array_equal(out[0, :-1, 2:-1, 0], x[0, :-1, :-3, 0])


def test_fp32_noise_floor_against_fp64(golden):
    """warp coordinates computed in fp64 from the same cameras: the fp32 oracle agrees to ~1e-5."""
    cams = golden['ex0_cams'].astype(np.float64)
    ds, di = float(golden['ds'][0]), float(golden['di'][0])

    def H64(l, r, d):
        Rl, Rr, tl, tr = l[0, :3, :3], r[0, :3, :3], l[0, :3, 3:], r[0, :3, 3:]
        Kl, Kr = l[1, :3, :3], r[1, :3, :3]
        cl, cr = -Rl.T @ tl, -Rr.T @ tr
        return Kr @ Rr @ (np.eye(3) - (cr - cl) @ Rl[2:3] * (ds + d * di)) @ Rl.T @ np.linalg.inv(Kl)
    H32 = ohw.get_homographies(golden['ex0_cams'][0:1], golden['ex0_cams'][1:2], 128, golden['ds'], golden['di'])
    for d in (0, 64, 127):
        h64 = H64(cams[0], cams[1], d)
        assert np.abs(H32[0, d] - h64).max() / np.abs(h64).max() < 2e-6


def test_tf_same_padding_semantics():
    # stride 2, k 3 on an even extent pads (0,1): out[i] = x[2i]*w0 + x[2i+1]*w1 + x[2i+2]*w2
    x = np.zeros((1, 1, 1, 6, 1), np.float32)
    x[0, 0, 0, :, 0] = [1, 2, 3, 4, 5, 6]
    w = np.zeros((3, 3, 3, 1, 1), np.float32)
    w[1, 1, :, 0, 0] = [1, 10, 100]
    y = onet.conv3d(x, w, 2)[0, 0, 0, :, 0]
    assert np.array_equal(y, [321, 543, 65])
    y1 = onet.conv3d(x, w, 1)[0, 0, 0, :, 0]                      # stride 1 pads (1,1)
    assert np.array_equal(y1, [210, 321, 432, 543, 654, 65])
    # conv3d_transpose SAME stride 2: out[2i+k] += in[i]*w[k], cropped to [0, 2n)
    xt = np.zeros((1, 1, 1, 3, 1), np.float32)
    xt[0, 0, 0, :, 0] = [1, 2, 3]
    wt = np.zeros((3, 3, 3, 1, 1), np.float32)
    wt[0, 0, :, 0, 0] = [1, 10, 100]           # only kd=kh=0 contributes to output (0,0,:)
    yt = onet.deconv3d(xt, wt)
    assert yt.shape == (1, 2, 2, 6, 1)
    assert np.array_equal(yt[0, 0, 0, :, 0], [1, 10, 102, 20, 203, 30])


def test_deconv_is_adjoint_of_strided_conv():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 4, 6, 8, 3)).astype(np.float64).astype(np.float32)
    y = rng.standard_normal((1, 2, 3, 4, 5)).astype(np.float32)
    w = rng.standard_normal((3, 3, 3, 3, 5)).astype(np.float32)     # conv kernel [.,.,.,Cin=3,Cout=5]
    lhs = float((onet.conv3d(x, w, 2).astype(np.float64) * y).sum())
    rhs = float((x.astype(np.float64) * onet.deconv3d(y, w)).sum())   # same array read as [.,.,.,Cout=3,Cin=5]
    assert abs(lhs - rhs) < 1e-3 * abs(lhs)


def test_batch_norm_train_statistics():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((1, 4, 5, 6, 3)) * [1, 5, 0.1] + [0, 3, -2]).astype(np.float32)
    y = onet.batch_norm_train(x)
    assert np.abs(y.mean(axis=(0, 1, 2, 3))).max() < 1e-5
    v = x.var(axis=(0, 1, 2, 3))
    assert np.allclose(y.var(axis=(0, 1, 2, 3)), v / (v + 1e-3), rtol=1e-4)


def test_prob_map_counts_integral_estimate_twice():
    D = 8
    v = np.full((1, D, 1, 1), 40.0, np.float32)
    v[0, 3] = -40.0                                                  # all mass on plane 3
    est, pm = om.prob2depth(v, D, np.float32([1.0]), np.float32([1.0]), out_prob_map=True)
    assert np.allclose(est, 4.0)
    assert np.allclose(pm, 2.0, atol=1e-6)                           # floor == ceil == 3 (model.py:42-62)
