"""CPU: known answers for the depth-map fusion oracle (oracle/fusibile.py, restating fusibile/fusibile.cu:138-325)."""
import numpy as np

from fusion_scene import make_scene
from oracle import fusibile as ofu


def _syn():
    import atvsnet_b200 as A
    return A.synthetic


def test_texture_fetch_restatement():
    rng = np.random.default_rng(1)
    img = rng.standard_normal((5, 7, 4)).astype(np.float32)
    xs, ys = np.meshgrid(np.arange(7, dtype=np.float32), np.arange(5, dtype=np.float32))
    assert np.array_equal(ofu.tex2d_linear(img, xs, ys), img)                 # texel centres are exact
    mid = ofu.tex2d_linear(img, np.float32([2.5]), np.float32([1.0]))
    assert np.allclose(mid[0], 0.5 * (img[1, 2] + img[1, 3]), rtol=1e-6)
    # 8-bit fractional weights: 1/512 rounds to 1/256, 1/1024 rounds to 0
    q = ofu.tex2d_linear(img, np.float32([2 + 1 / 512.0, 2 + 1 / 1024.0]), np.float32([0, 0]))
    assert np.allclose(q[0], img[0, 2] + (img[0, 3] - img[0, 2]) / 256.0, rtol=1e-5, atol=1e-6) and np.array_equal(q[1], img[0, 2])
    edge = ofu.tex2d_linear(img, np.float32([6.75]), np.float32([4.5]))       # clamp: beyond the last texel
    assert np.allclose(edge[0], img[4, 6])


def test_consistent_plane_is_kept_and_outliers_are_dropped():
    K, R, t, depths, images = make_scene(_syn(), outliers=False)
    cams = [ofu.camera_from_krt(K[i], R[i], t[i]) for i in range(len(K))]
    nd = np.concatenate([np.stack([ofu.fake_normals(d) for d in depths]), depths[..., None]], axis=-1)
    keep, X, n, tex, cnt = ofu.fuse_reference(0, nd, images, cams, 0.01, np.deg2rad(360.0), 2)
    # every pixel of the reference whose point is seen by >= 2 other views is kept, and lies on the plane z = 5
    assert keep.mean() > 0.8 and np.abs(X[keep][:, 2] - 5.0).max() < 1e-3
    assert np.allclose(n[keep][:, :3], 1 / 1.732050808, rtol=1e-5)
    assert cnt.max() == len(K) - 1
    # a block of view 1 scaled by 1.5: those reference pixels of view 1 are not confirmed by anybody
    K, R, t, depths, images = make_scene(_syn(), outliers=True)
    nd = np.concatenate([np.stack([ofu.fake_normals(d) for d in depths]), depths[..., None]], axis=-1)
    keep1, _, _, _, cnt1 = ofu.fuse_reference(1, nd, images, cams, 0.01, np.deg2rad(360.0), 2)
    assert not keep1[11:19, 14:28].any() and keep1.mean() > 0.5
    pts, nrm, tex = ofu.fuse(nd, images, cams, 0.01, np.deg2rad(360.0), 2)
    assert pts.shape[1] == 3 and nrm.shape == pts.shape and tex.shape == (pts.shape[0], 4)
    assert np.abs(pts[:, 2] - 5.0).max() < 0.05
    # stricter consensus keeps fewer points
    pts3, _, _ = ofu.fuse(nd, images, cams, 0.01, np.deg2rad(360.0), 3)
    assert 0 < pts3.shape[0] < pts.shape[0]
