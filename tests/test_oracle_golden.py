"""CPU: the oracle against tests/golden/reference_golden.npz, i.e. against outputs of the
reference's own source files executed on the NumPy TF stand-in (tests/golden/make_golden.py)."""
import numpy as np

from conftest import rel_err
from oracle import homography_warping as ohw
from oracle import model as om


def test_get_homographies_example_cams(golden):
    cams, ds, di = golden['ex0_cams'], golden['ds'], golden['di']
    for v in (1, 4):
        H = ohw.get_homographies(cams[0:1], cams[v:v + 1], 128, ds, di)
        assert H.shape == (1, 128, 3, 3)
        # the stand-in inverts K with LAPACK LU, the oracle with cofactors: fp32-level agreement
        assert rel_err(H, golden['ex0_H_0_%d' % v]) < 5e-6
    H = ohw.get_homographies(cams[3:4], cams[0:1], 16, ds, di)
    assert rel_err(H, golden['ex0_H_3_0']) < 5e-6
    H = ohw.get_homographies(cams[0:1], cams[2:3], 8, np.float32([2.0]), np.float32([0.5]), inverse_depth=False)
    assert rel_err(H, golden['ex0_H_0_2_depth']) < 5e-6


def test_homography_warping_bit_exact_given_H(golden):
    img, Hs = golden['warp_img'], golden['warp_H']
    for d in (0, 3, 7):
        out, mask = ohw.homography_warping(img, Hs[:, d], output_mask=True)
        assert np.array_equal(mask, golden['warp_mask_%d' % d])
        assert np.array_equal(out, golden['warp_bilinear_%d' % d])      # same fp32 op order
        assert 0.05 < mask.mean() < 1.0
    out, mask = ohw.homography_warping(img[..., :1], Hs[:, 3], method='nearest', output_mask=True)
    assert np.array_equal(mask, golden['warp_nearest_mask_3'])
    assert np.array_equal(out, golden['warp_nearest_3'])


def test_homography_warping_by_depth(golden):
    c = golden['warp_cams']
    out, mask = ohw.homography_warping_by_depth(golden['warp_img'], c[0:1], c[1:2], golden['bydepth_depth'],
                                                output_mask=True)
    g = golden['bydepth_out']
    same = mask == golden['bydepth_mask']
    assert same.mean() > 0.999            # K^-1 differs in the last bit: a border pixel may flip
    ok = same[..., 0]
    assert np.abs(out[ok] - g[ok]).max() < 2e-4 * np.abs(g).max()


def test_build_cost_volume(golden):
    c = golden['warp_cams'][None, :2]
    ds, di = golden['ds'], golden['di'] * 32
    ref, view = golden['cv_ref'], golden['warp_img']
    cv = om.build_cost_volume(ref, view, c, 4, ds, di, 0, 1)
    g = golden['cv_concat']
    assert cv.shape == g.shape == (1, 4, 40, 60, 16)
    assert np.array_equal(cv[..., :8], g[..., :8])                       # tiled reference half
    bad = np.abs(cv - g).max(axis=-1) > 2e-4 * np.abs(g).max()
    assert bad.mean() < 1e-3
    cv = om.build_cost_volume(view, ref, c, 4, ds, di, 1, 0)
    bad = np.abs(cv - golden['cv_concat_rev']).max(axis=-1) > 2e-4 * np.abs(g).max()
    assert bad.mean() < 1e-3
    cv = om.build_cost_volume(ref, view, c, 2, ds, di, 0, 1, warp_ref=True)
    bad = np.abs(cv - golden['cv_warpref']).max(axis=-1) > 2e-4 * np.abs(g).max()
    assert bad.mean() < 1e-3


def test_cost_volume_reasoning(golden, gweights):
    prob, filt = om.cost_volume_reasoning(golden['crm_in'], gweights, output_filtered_cost=True)
    assert rel_err(filt, golden['crm_filtered']) < 2e-4
    assert rel_err(prob, golden['crm_prob']) < 2e-4
    only = om.cost_volume_reasoning(golden['crm_in'], gweights, output_prob=False)
    assert rel_err(only, golden['crm_filtered_only']) < 2e-4


def test_crm_intermediate_layers(golden, gweights):
    from oracle import network as onet
    tower = onet.StackedUNet_prob({'data': golden['crm_in']}, gweights)
    for nm in ('conv_b0_1_0', 'conv_b0_0_1', 'conv_b0_3_1', 'conv_b0_4_0', 'conv_b0_6_0', 'conv_b1_0_0',
               'conv_b1_5_0'):
        assert rel_err(tower.get_output_by_name(nm), golden['crm_' + nm]) < 2e-4, nm


def test_attention_aggregation_and_output_conv(golden, gweights):
    xs = golden['aam_in']
    keep = om.cost_volume_aggregation(xs, gweights, keepchannel=True)
    assert rel_err(keep, golden['aam1_keep']) < 1e-5
    assert rel_err(om.cost_volume_aggregation(xs, gweights, keepchannel=False), golden['aam1_prob']) < 1e-5
    assert rel_err(om.cost_volume_aggregation_refine(xs, gweights, keepchannel=True), golden['aam2_keep']) < 1e-5
    assert rel_err(om.output_conv(golden['aam1_keep'], gweights), golden['outconv']) < 1e-5
    assert rel_err(om.output_conv_refine(golden['aam1_keep'], gweights), golden['outconv_refine']) < 1e-5


def test_prob2depth(golden):
    vol, ds, di = golden['p2d_vol'], golden['p2d_start'], golden['p2d_interval']
    est, pm = om.prob2depth(vol, 16, ds, di, out_prob_map=True)
    assert rel_err(est, golden['p2d_est']) < 1e-6
    assert np.abs(pm - golden['p2d_prob']).max() < 1e-6
    e, eu, p, pu = om.prob2depth_upsample(vol, 16, ds, di, out_prob_map=True)
    assert eu.shape == (1, 40, 48, 1)
    assert rel_err(eu, golden['p2d_est_up']) < 1e-6
    assert (np.abs(pu - golden['p2d_prob_up']) > 1e-5).mean() < 1e-3
