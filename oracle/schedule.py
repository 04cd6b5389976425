"""The reference's two drivers as CPU schedules over the oracle pieces, IMAGES in (TEST INFRASTRUCTURE, see
oracle/__init__.py):

  run_multiview  /root/reference/atvsnet/example.py:144-191 (run_test_multiview): FEM per view, stage I
                 (TVSNet_base_siamese per source), stage II (AAM1 -> output_conv -> prob2depth), stage III
                 (TVSNet_refine per source), stage IV (AAM2 -> output_conv_refine -> prob2depth_upsample) and the
                 host epilogue ``out[out < 1e-10] = inf; depth = 1 / out``
  run_twoview    /root/reference/atvsnet/example.py:219-272 (run_test_twoview): TVSNet (model.py:346-377) ->
                 prob2depth_upsample, host epilogue ``out[out <= 0] = inf; depth = 1 / out``

The reference recomputes the FEM of the reference image in every sess.run; the towers share weights and see the same
input, so computing each view's features once gives the same numbers."""
import numpy as np

from . import fem
from . import model as om
from . import refine as oref

F32 = np.float32


def extract_features(images, weights):
    """model.py:420-425 over all views: (B,N,H,W,3) -> (B,N,H/4,W/4,32); one tower per view (statistics over B)."""
    images = np.asarray(images, dtype=F32)
    return np.stack([fem.ResNetDS2SPP(images[:, n], weights) for n in range(images.shape[1])], axis=1)


def inverse_to_depth(out, twoview=False):
    """example.py:183-186 (multi-view: values below 1e-10 -> inf) / :269-272 (two-view: values <= 0 -> inf)."""
    out = np.array(out, dtype=F32, copy=True)
    out[(out <= 0) if twoview else (out < 1e-10)] = np.inf
    return (F32(1.0) / out).astype(F32)


def run_multiview(images, cams, depth_num, weights, features=None, refine=True):
    """example.py:144-191.  images (B,N,H,W,3) raw 0..255 BGR, cams (B,N,2,4,4) at feature resolution."""
    cams = np.asarray(cams, dtype=F32)
    images = np.asarray(images, dtype=F32)
    ds, di = cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1]
    feats = features if features is not None else extract_features(images, weights)
    out = om.run_multiview_stage12(feats, cams, depth_num, weights, siamese=True)
    out['features'] = feats
    if not refine:
        return out
    rcs, rps = [], []
    for n, v in enumerate(range(1, cams.shape[1])):
        rp, rc = oref.TVSNet_refine(out['depth_agg_init'], out['depth_views'][n], out['prob_volume_agg'],
                                    out['cost_volume_agg'], images, cams, depth_num, ds, di, v, weights)
        rps.append(rp)
        rcs.append(rc)
    cref = om.cost_volume_aggregation_refine(np.stack(rcs, axis=-1), weights, keepchannel=True)
    pref = om.output_conv_refine(cref, weights)
    est, est_up = om.prob2depth_upsample(pref, depth_num, ds, di)
    out.update(refined_cost_volume_agg=cref, refined_prob_volume_agg=pref, depth_refined=est, depth_refined_up=est_up,
               pred=inverse_to_depth(est_up))
    return out


def TVSNet(images, cams, depth_num, depth_start, depth_interval, view_i, weights, ref_i=0):
    """model.py:346-377 -> refined_prob_vol (B,D,h,w) (+ the intermediate stage outputs)."""
    images = np.asarray(images, dtype=F32)
    ref = fem.ResNetDS2SPP(images[:, ref_i], weights)
    view = fem.ResNetDS2SPP(images[:, view_i], weights)
    cv_view = om.build_cost_volume(view, ref, cams, depth_num, depth_start, depth_interval, ref_id=view_i, view_id=0)
    depth_view = om.prob2depth(om.cost_volume_reasoning(cv_view, weights), depth_num, depth_start, depth_interval)
    cv = om.build_cost_volume(ref, view, cams, depth_num, depth_start, depth_interval, ref_id=0, view_id=view_i)
    prob, _ = om.cost_volume_reasoning(cv, weights, output_filtered_cost=True)
    depth_b2 = om.prob2depth(prob, depth_num, depth_start, depth_interval)
    init = np.stack([depth_b2, depth_view], axis=1)
    _, prob_res = oref.refinement(init, cams, depth_num, depth_start, depth_interval, images, prob, ref_i, view_i, weights,
                                  num_depths=2, depth_ref_id=0, depth_view_id=1)
    return prob + prob_res, dict(depth_b2=depth_b2, depth_view=depth_view, prob_vol_b2=prob)


def run_twoview(images, cams, depth_num, weights):
    """example.py:219-272."""
    cams = np.asarray(cams, dtype=F32)
    ds, di = cams[:, 0, 1, 3, 0], cams[:, 0, 1, 3, 1]
    refined, mid = TVSNet(images, cams, depth_num, ds, di, 1, weights, 0)
    est, est_up = om.prob2depth_upsample(refined, depth_num, ds, di)
    mid.update(refined_prob_volume=refined, depth_refined=est, depth_refined_up=est_up,
               pred=inverse_to_depth(est_up, twoview=True))
    return mid
