"""GPU: refinement stage (a-tvsnet_b200/refine.py through the C ABI, fp32 first path) against the golden vectors produced by
the reference's own functions on the TF stand-in (tests/golden/reference_golden_refine.npz) and the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope='module')
def A():
    import atvsnet_b200 as A_
    return A_


@pytest.fixture(scope='module')
def g():
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden_refine.npz')))


def test_transform_depth_and_visual_hull(A, g):
    cams, v = g['cams'], int(g['view_i'])
    td = A.refine.transform_depth(cu(g['depth_view']), cu(cams[:, v]), cu(cams[:, 0])).cpu().numpy()
    assert rel(td, g['transform_depth']) < 1e-5
    assert np.all(td[0, 0, :3] == 0)
    ds, di = cu(cams[:, 0, 1, 3, 0]), cu(cams[:, 0, 1, 3, 1])
    depths = np.stack([g['depth_b2'], g['depth_view']], 1)[..., 0]
    vh = A.refine.get_visual_hull(cu(depths), cu(cams), 8, ds, di, ref_id=0, view_num=2).cpu().numpy()
    assert vh.shape == g['visual_hull'].shape
    assert (np.abs(vh - g['visual_hull']) > 1e-6).mean() < 2e-3
    assert set(np.unique(vh)).issubset({0.0, 0.5, 1.0})


def test_refinement_matches_reference_functions(A, g):
    from gen_common import named_weights
    from oracle import refine as oref
    w = named_weights('refine_variables.json', 5)
    A.variables.load_weights(w)
    cams, v = g['cams'], int(g['view_i'])
    imgs = cu(g['images'])
    rf, vf = A.refine.extract_feature_shallow(imgs, 0, v)
    assert rel(rf.cpu().numpy(), g['shallow_ref']) < 1e-4 and rel(vf.cpu().numpy(), g['shallow_view']) < 1e-4
    ds, di = cu(cams[:, 0, 1, 3, 0]), cu(cams[:, 0, 1, 3, 1])
    # the two hand-assembled input groups against the oracle's
    init = np.stack([g['depth_b2'], g['depth_view']], axis=1)
    _, _, grp = oref.refinement(init, cams, 8, cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1], g['images'], g['prob'], 0, v, w,
                                num_depths=2, depth_ref_id=0, depth_view_id=1, return_groups=True)
    rp, rc = A.refine.TVSNet_refine(cu(g['depth_b2']), cu(g['depth_view']), cu(g['prob']), cu(g['cost']), imgs, cu(cams), 8,
                                    ds, di, v)
    torch.cuda.synchronize()
    assert tuple(rp.shape) == g['refined_prob'].shape and tuple(rc.shape) == g['refined_cost'].shape
    assert rel(rp.cpu().numpy() - g['prob'], g['refined_prob'] - g['prob']) < 2e-4
    assert rel(rc.cpu().numpy() - g['cost'], g['refined_cost'] - g['cost']) < 2e-4
    assert grp['geo_group'].shape[-1] == 19 and grp['photo_group'].shape[-1] == 48
