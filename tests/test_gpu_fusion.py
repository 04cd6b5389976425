"""GPU: depth-map fusion (csrc/fusion.cu through atvs_fuse_depth_maps) against the CPU oracle of the fusibile kernel and
its host scan (oracle/fusibile.py, fusibile/fusibile.cu:138-325): same points in the same order."""
import numpy as np
import pytest
import torch

from fusion_scene import make_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def A():
    import atvsnet_b200 as A_
    return A_


@pytest.mark.parametrize('n_views,H,W,ncons', [(4, 48, 64, 2), (5, 37, 53, 2), (3, 64, 96, 1), (4, 48, 64, 3)])
def test_fusion_matches_oracle(A, n_views, H, W, ncons):
    from oracle import fusibile as ofu
    K, R, t, depths, images = make_scene(A.synthetic, n_views, H, W, seed=n_views)
    cams_o = [ofu.camera_from_krt(K[i], R[i], t[i]) for i in range(n_views)]
    nd = np.concatenate([np.stack([ofu.fake_normals(d) for d in depths]), depths[..., None]], axis=-1)
    pts, nrm, tex = ofu.fuse(nd, images, cams_o, 0.01, np.deg2rad(360.0), ncons)
    cams = A.fusion.cameras_from_krt(K, R, t)
    out = A.fusion.fuse_depth_maps(torch.from_numpy(depths).cuda(), cams, images=torch.from_numpy(images).cuda(),
                                   disp_thresh=0.01, num_consistent=ncons)
    assert out['count'] == pts.shape[0]
    assert np.array_equal(out['points'].cpu().numpy(), pts)                # bit-exact coordinates, reference order
    assert np.allclose(out['normals'].cpu().numpy(), nrm, rtol=1e-6, atol=1e-7)
    assert np.allclose(out['colors'].cpu().numpy(), tex, rtol=1e-5, atol=1e-4)
    # without colours, with explicit normals, and with a capacity smaller than the cloud
    out2 = A.fusion.fuse_depth_maps(torch.from_numpy(depths).cuda(), cams, normals=A.fusion.fake_normals(torch.from_numpy(depths).cuda()),
                                    disp_thresh=0.01, num_consistent=ncons, capacity=100)
    assert out2['count'] == pts.shape[0] and out2['points'].shape[0] == min(100, pts.shape[0]) and out2['colors'] is None
    assert np.array_equal(out2['points'].cpu().numpy(), pts[:100])


def test_probability_filter_and_ply(A, tmp_path):
    d = torch.rand(2, 8, 8, device='cuda') + 1
    p = torch.rand(2, 8, 8, device='cuda')
    f = A.fusion.probability_filter(d, p, 0.8)
    assert bool(((f == 0) == (p < 0.8)).all()) and bool((f[p >= 0.8] == d[p >= 0.8]).all())
    pts = torch.rand(10, 3, device='cuda')
    A.fusion.write_ply(str(tmp_path / 'c.ply'), pts, pts, torch.rand(10, 4, device='cuda') * 255)
    raw = open(str(tmp_path / 'c.ply'), 'rb').read()
    assert raw.startswith(b'ply\nformat binary_little_endian 1.0\nelement vertex 10\n') and len(raw.split(b'end_header\n', 1)[1]) == 10 * 27
