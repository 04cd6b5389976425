"""CPU: the refinement-stage oracle (oracle/refine.py, SURVEY.md 8(f) row N2 - checker written ahead of the CUDA code)
against the reference's own refinement functions executed on the TF stand-in (tests/golden/make_golden_refine.py)."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from gen_common import named_weights  # noqa: E402
from oracle import refine  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='module')
def g():
    return dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden_refine.npz')))


def test_refine_variable_list():
    shapes = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'refine_variables.json')))
    assert len(shapes) == 39
    assert shapes['global_refine_geo_3dconv/conv3d/kernel'] == [3, 3, 3, 19, 8]       # 1 + 16 + 1 + 1
    assert shapes['global_refine_photo_3dconv/conv3d/kernel'] == [3, 3, 3, 48, 8]
    assert shapes['global_refine_conv0_x_0/shortcut/weights'] == [1, 1, 3, 16]
    assert shapes['global_refine_3dconv4_0/conv3d_transpose/kernel'] == [3, 3, 3, 32, 64]
    assert shapes['global_refined_cost_vol/kernel'] == [3, 3, 3, 8, 1]


def test_transform_depth_and_visual_hull(g):
    cams, v = g['cams'], int(g['view_i'])
    td = refine.transform_depth(g['depth_view'], cams[:, v], cams[:, 0])
    assert rel(td, g['transform_depth']) < 1e-5
    assert np.all(td[0, 0, :3] == 0)                                   # zero inverse depth stays invalid (masked)
    ds, di = cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1]
    vh = refine.get_visual_hull(np.stack([g['depth_b2'], g['depth_view']], 1)[..., 0], cams, 8, ds, di, ref_id=0, view_num=2)
    assert vh.shape == g['visual_hull'].shape == (1, 8, 8, 16, 1)
    assert (np.abs(vh - g['visual_hull']) > 1e-6).mean() < 2e-3        # counts in {0, .5, 1}: a flip needs a tie
    # a depth image in front of every plane is "behind" none of them, one behind every plane is behind all
    near = np.full((1, 2, 8, 16), 10.0, np.float32)
    assert np.all(refine.get_visual_hull(near, cams, 8, ds, di, view_num=2)[..., 0][:, :, 2:6, 4:12] >= 0.5)


def test_shallow_features_and_refinement(g):
    w = named_weights('refine_variables.json', 5)
    v = int(g['view_i'])
    assert rel(refine.shallow_features(g['images'][:, 0], w), g['shallow_ref']) < 2e-5
    assert rel(refine.shallow_features(g['images'][:, v], w), g['shallow_view']) < 2e-5
    cams = g['cams']
    ds, di = cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1]
    rp, rc = refine.TVSNet_refine(g['depth_b2'], g['depth_view'], g['prob'], g['cost'], g['images'], cams, 8, ds, di, v, w)
    assert rp.shape == g['refined_prob'].shape and rc.shape == g['refined_cost'].shape
    # residuals of the refinement net on top of the inputs: compare the residual parts
    assert rel(rp - g["prob"], g["refined_prob"] - g["prob"]) < 5e-5
    assert rel(rc - g["cost"], g["refined_cost"] - g["cost"]) < 5e-5
