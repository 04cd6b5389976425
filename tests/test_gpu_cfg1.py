"""GPU, BASELINE.json configs[0]: atvsnet/example.py on the bundled scenes with 3 views - the REAL images and cameras
(tests/golden/example/, copied from the reference's example/ by make_example_fixtures.py), seeded synthetic weights
under the checkpoint names (the released model.zip is not available offline), CUDA path against the
same schedule composed from the CPU oracles (oracle/schedule.py):

  example/0  3 views, 960x640 images -> 160x240 features, D = 128: the four-stage multi-view schedule
             (example.py:144-191), images in, refinement included;
  example/2  2 views, 640x480: the two-view network TVSNet + prob2depth_upsample (example.py:219-272).

Each test runs ~1-2 minutes of CPU oracle next to the GPU run (pytest -m "gpu and not slow" skips them)."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, 'tests', 'golden', 'example')


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def npy(t):
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope='module')
def A():
    import atvsnet_b200 as A_
    return A_


def all_weights(A):
    w = A.variables.synthetic_weights()          # bench.py's CRM / AAM / output-conv weights (logit gain 4)
    w.update(A.variables.synthetic_fem_weights())
    w.update(A.variables.synthetic_refine_weights())
    return w


def mae_over_range(a, b, cams, D):
    return float(np.abs(a - b).mean()) / ((D - 1) * float(cams[0, 0, 1, 3, 1]))


def test_example0_multiview_3_views(A):
    from oracle import schedule as osch
    D = 128
    images, cams, _ = A.pipeline.load_example(os.path.join(EX, '0'), view_num=3)
    assert images.shape == (1, 3, 640, 960, 3) and cams.shape == (1, 3, 2, 4, 4)
    w = all_weights(A)
    A.variables.load_weights(w)
    assert A.FLAGS.precision == 'fp16'
    out = A.pipeline.run_example(cu(images), cu(cams), D)
    torch.cuda.synchronize()
    ref = osch.run_multiview(images, cams, D, w)
    m2 = mae_over_range(npy(out['depth']), ref['depth_agg_init'], cams, D)
    m4 = mae_over_range(npy(out['depth_refined_up']), ref['depth_refined_up'], cams, D)
    p = torch.softmax(-torch.from_numpy(ref['refined_prob_volume_agg']), dim=1).max(dim=1).values.mean().item()
    print("example/0, 3 views: stage II depth MAE / range %.3e, final (stage IV, x4) %.3e, mean peak probability %.3f"
          % (m2, m4, p))
    assert tuple(out['depth_refined_up'].shape) == (1, 640, 960, 1)
    assert m2 < 1e-3 and m4 < 1e-3, (m2, m4)
    # host epilogue (example.py:183-186): depth = 1 / inverse depth, inside the swept range
    pred = npy(out['pred'])
    lo, hi = float(cams[0, 0, 1, 3, 0]), float(cams[0, 0, 1, 3, 0] + (D - 1) * cams[0, 0, 1, 3, 1])
    assert np.isfinite(pred).all() and pred.min() >= 1.0 / hi - 1e-4 and pred.max() <= 1.0 / lo + 1e-4
    assert np.abs(pred - ref['pred']).mean() / np.abs(ref['pred']).mean() < 2e-3


def test_example2_twoview(A):
    from oracle import schedule as osch
    D = 128
    images, cams, _ = A.pipeline.load_example(os.path.join(EX, '2'), view_num=3)     # only 2 views exist -> two-view path
    assert images.shape == (1, 2, 480, 640, 3)
    w = all_weights(A)
    A.variables.load_weights(w)
    out = A.pipeline.run_example(cu(images), cu(cams), D)
    torch.cuda.synchronize()
    ref = osch.run_twoview(images, cams, D, w)
    m = mae_over_range(npy(out['depth_refined_up']), ref['depth_refined_up'], cams, D)
    print("example/2, two-view: refined depth MAE / range %.3e" % m)
    assert tuple(out['depth_refined_up'].shape) == (1, 480, 640, 1)
    assert m < 1e-3, m
    assert np.abs(npy(out['pred']) - ref['pred']).mean() / np.abs(ref['pred']).mean() < 2e-3
