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
    # fp32 CUDA-core path: fp32 summation order only; fp16 tensor-core path (the default): 16 layers of 11-bit storage on
    # an 8 x 16 x 24 volume whose deepest level has 6 voxels per channel for its batch statistics
    for prec, tol, tol_mean in (('fp32', 2e-4, 2e-4), ('fp16', 0.25, 0.02)):
        A.FLAGS.precision = prec
        try:
            rp, rc = A.refine.TVSNet_refine(cu(g['depth_b2']), cu(g['depth_view']), cu(g['prob']), cu(g['cost']), imgs,
                                            cu(cams), 8, ds, di, v)
            torch.cuda.synchronize()
        finally:
            A.FLAGS.precision = A.flags.DEFAULT_PRECISION
        assert tuple(rp.shape) == g['refined_prob'].shape and tuple(rc.shape) == g['refined_cost'].shape
        dp, dc = g['refined_prob'] - g['prob'], g['refined_cost'] - g['cost']
        ep, ec = np.abs(rp.cpu().numpy() - g['prob'] - dp), np.abs(rc.cpu().numpy() - g['cost'] - dc)
        print("TVSNet_refine %s: residual max rel err prob %.2e cost %.2e, mean rel err prob %.2e cost %.2e"
              % (prec, ep.max() / np.abs(dp).max(), ec.max() / np.abs(dc).max(), ep.mean() / np.abs(dp).mean(),
                 ec.mean() / np.abs(dc).mean()))
        assert ep.max() / np.abs(dp).max() < tol and ec.max() / np.abs(dc).max() < tol
        assert ep.mean() / np.abs(dp).mean() < tol_mean and ec.mean() / np.abs(dc).mean() < tol_mean
    assert grp['geo_group'].shape[-1] == 19 and grp['photo_group'].shape[-1] == 48


def test_four_stage_schedule_fp32_against_oracle(A):
    """example.py:144-181 end to end, images in: FEM -> stage I/II -> TVSNet_refine per source -> AAM2 -> x4 soft-argmin,
    fp32 CUDA path against the same schedule composed from the CPU oracles."""
    from gen_common import fem_weights, named_weights
    from oracle import fem as ofem
    from oracle import model as om
    from oracle import refine as oref
    rng = np.random.default_rng(17)
    nv, H, W, D = 3, 32, 64, 8
    h, w = H // 4, W // 4
    imgs = (127.5 + 50 * rng.standard_normal((1, nv, H, W, 3))).clip(0, 255).astype(np.float32)
    cams = A.synthetic.orbit_cams(nv, h, w, D)[None]
    ds, di = cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1]
    allw = A.variables.synthetic_weights(seed=11, logit_gain=2.0)
    allw.update(fem_weights(7))
    allw.update(named_weights('refine_variables.json', 5))
    # ---- oracle schedule
    feats = np.stack([ofem.ResNetDS2SPP(imgs[:, n], allw)[0] for n in range(nv)])[None]
    s12 = om.run_multiview_stage12(feats, cams, D, allw, siamese=True)
    rcs = []
    for n, v in enumerate(range(1, nv)):
        rp, rc = oref.TVSNet_refine(s12['depth_agg_init'], s12['depth_views'][n], s12['prob_volume_agg'],
                                    s12['cost_volume_agg'], imgs, cams, D, ds, di, v, allw)
        rcs.append(rc)
    cref = om.cost_volume_aggregation_refine(np.stack(rcs, axis=-1), allw, keepchannel=True)
    pref = om.output_conv_refine(cref, allw)
    _, est_up = om.prob2depth_upsample(pref, D, ds, di)
    # ---- CUDA schedule
    A.variables.load_weights(allw)
    A.FLAGS.precision = 'fp32'
    try:
        out = A.pipeline.run_example_schedule(cu(imgs), cu(cams), D)
        torch.cuda.synchronize()
    finally:
        A.FLAGS.precision = A.flags.DEFAULT_PRECISION
    rng_ = float((D - 1) * di[0])
    assert tuple(out['depth_refined_up'].shape) == est_up.shape
    assert float(np.abs(out['depth'].cpu().numpy() - s12['depth_agg_init']).mean()) / rng_ < 1e-3
    assert float(np.abs(out['depth_refined_up'].cpu().numpy() - est_up).mean()) / rng_ < 2e-3
