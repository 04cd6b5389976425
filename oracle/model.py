"""NumPy fp32 restatement of /root/reference/atvsnet/model.py for the hot path
(get_propability_map :13-65, upsample_prob_vol :68-76, prob2depth :80-109,
prob2depth_upsample :113-129, output_conv :132-140, build_cost_volume :157-200,
cost_volume_reasoning :204-223, TVSNet_base :380-395, TVSNet_base_siamese :398-417,
cost_volume_aggregation(_refine) :445-468) plus the stage-I/II schedule of
/root/reference/atvsnet/example.py:144-158.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Same function names, argument order
and layouts as the reference; ``weights`` (dict keyed by TF variable names) and
``inverse_depth`` replace the TF variable store and ``FLAGS``.  The FEM
(ResNetDS2SPP) is outside the current scope, so the TVSNet_* entry points take
*features* (B,N,h,w,F) where the reference takes images.
"""
import numpy as np

from . import network as net
from .homography_warping import get_homographies, homography_warping

F32 = np.float32


def _f(x):
    return np.asarray(x, dtype=F32)


def get_propability_map(cv, depth_map, depth_start, depth_interval):
    """model.py:13-65.  cv = softmax volume (B,D,H,W); depth_map (B,H,W,1).
    (The reference's index construction is only right for B == 1; this restates the
    intended per-sample gather.)"""
    cv = _f(cv)
    B, D, H, W = cv.shape
    ds = _f(depth_start).reshape(B, 1, 1, 1)
    di = _f(depth_interval).reshape(B, 1, 1, 1)
    d = ((_f(depth_map) - ds) / di)[..., 0]                      # (B,H,W)
    l0 = np.clip(np.floor(d).astype(np.int32), 0, D - 1)
    l1 = np.clip(l0 - 1, 0, D - 1)
    r0 = np.clip(np.ceil(d).astype(np.int32), 0, D - 1)
    r1 = np.clip(r0 + 1, 0, D - 1)
    b, y, x = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing='ij')
    prob = ((cv[b, l0, y, x] + cv[b, l1, y, x]) + cv[b, r0, y, x]) + cv[b, r1, y, x]
    return prob[..., None]


def upsample_prob_vol(prob_vol, up_scale=4):
    """model.py:68-76: bilinear, align_corners=True, on every depth slice."""
    v = _f(prob_vol)
    B, D, H, W = v.shape
    Ho, Wo = H * up_scale, W * up_scale

    def axis(n_in, n_out):
        scale = F32(n_in - 1) / F32(n_out - 1) if n_out > 1 else F32(0)
        src = np.arange(n_out, dtype=F32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        return lo, hi, (src - lo.astype(F32)).astype(F32)

    y0, y1, fy = axis(H, Ho)
    x0, x1, fx = axis(W, Wo)
    fy = fy[None, None, :, None]
    fx = fx[None, None, None, :]
    top = v[:, :, y0][:, :, :, x0] + (v[:, :, y0][:, :, :, x1] - v[:, :, y0][:, :, :, x0]) * fx
    bot = v[:, :, y1][:, :, :, x0] + (v[:, :, y1][:, :, :, x1] - v[:, :, y1][:, :, :, x0]) * fx
    return (top + (bot - top) * fy).astype(F32)


def prob2depth(prob_volume, depth_num, depth_start, depth_interval, out_prob_map=False):
    """model.py:80-109.  (B,D,H,W) -> (B,H,W,1) [, prob map (B,H,W,1)]."""
    v = _f(prob_volume)
    B = v.shape[0]
    ds = _f(depth_start).reshape(B)
    di = _f(depth_interval).reshape(B)
    de = ds + (F32(depth_num) - F32(1)) * di
    neg = -v
    m = neg.max(axis=1, keepdims=True)
    e = np.exp(neg - m)
    p = e / e.sum(axis=1, keepdims=True, dtype=F32)
    # tf.linspace(start, stop, num): start + i * (stop - start) / (num - 1)
    step = (de - ds) / F32(max(depth_num - 1, 1))
    soft = ds[:, None] + np.arange(depth_num, dtype=F32)[None, :] * step[:, None]     # (B,D)
    est = (soft[:, :, None, None] * p).sum(axis=1, dtype=F32)[..., None]
    if out_prob_map:
        return est, get_propability_map(p, est, ds, di)
    return est


def prob2depth_upsample(prob_volume, depth_num, depth_start, depth_interval, out_prob_map=False):
    """model.py:113-129."""
    up = upsample_prob_vol(prob_volume)
    if out_prob_map:
        est_up, pm_up = prob2depth(up, depth_num, depth_start, depth_interval, True)
        est, pm = prob2depth(prob_volume, depth_num, depth_start, depth_interval, True)
        return est, est_up, pm, pm_up
    return (prob2depth(prob_volume, depth_num, depth_start, depth_interval),
            prob2depth(up, depth_num, depth_start, depth_interval))


def output_conv(cost_volume, weights, reuse=None):
    """model.py:132-135."""
    return net.OutputConv({'data': _f(cost_volume)}, weights).get_output()[..., 0]


def output_conv_refine(cost_volume, weights, reuse=None):
    """model.py:137-140."""
    return net.OutputConv_refine({'data': _f(cost_volume)}, weights).get_output()[..., 0]


def build_cost_volume(ref_feature, view_feature, cams, depth_num, depth_start, depth_interval,
                      ref_id, view_id, output_homo=False, warp_ref=False, inverse_depth=True):
    """model.py:157-200.  -> (B,D,h,w,2F) [, (B,D,3,3)]."""
    ref_feature, view_feature, cams = _f(ref_feature), _f(view_feature), _f(cams)
    ref_cam = cams[:, ref_id]
    view_cam = cams[:, view_id]
    hs = get_homographies(ref_cam, view_cam, depth_num, depth_start, depth_interval, inverse_depth)
    if warp_ref:
        rh = get_homographies(ref_cam, ref_cam, depth_num, depth_start, depth_interval, inverse_depth)
        cost = np.stack([homography_warping(ref_feature, rh[:, d]) for d in range(depth_num)], axis=1)
    else:
        cost = np.tile(ref_feature[:, None], (1, depth_num, 1, 1, 1))
    warped = np.stack([homography_warping(view_feature, hs[:, d]) for d in range(depth_num)], axis=1)
    cost = np.concatenate([cost, warped], axis=-1)
    return (cost, hs) if output_homo else cost


def cost_volume_reasoning(cost_volume, weights, output_prob=True, output_filtered_cost=False, reuse=None):
    """model.py:204-223."""
    if output_prob:
        tower = net.StackedUNet_prob({'data': _f(cost_volume)}, weights)
        prob = tower.get_output()[..., 0]
        if output_filtered_cost:
            return prob, tower.get_output_by_name('conv_b2_6_1')
        return prob
    tower = net.StackedUNet({'data': _f(cost_volume)}, weights)
    return tower.get_output_by_name('conv_b2_6_1')


def cost_volume_aggregation(cost_volumes, weights, reuse=None, keepchannel=False):
    """model.py:445-456.  (B,D,h,w,C,N-1) -> (B,D,h,w,C) | (B,D,h,w)."""
    if keepchannel:
        return net.AttAggregation_keepchannel({'data': _f(cost_volumes)}, weights).get_output()
    return net.AttAggregation({'data': _f(cost_volumes)}, weights).get_output()[..., 0]


def cost_volume_aggregation_refine(cost_volumes, weights, reuse=None, keepchannel=False):
    """model.py:460-468."""
    if keepchannel:
        return net.AttAggregation_refine_keepchannel({'data': _f(cost_volumes)}, weights).get_output()
    return net.AttAggregation_refine({'data': _f(cost_volumes)}, weights).get_output()[..., 0]


def TVSNet_base(features, cams, depth_num, depth_start, depth_interval, view_i, weights, ref_i=0,
                inverse_depth=True):
    """model.py:380-395 with features (B,N,h,w,F) in place of images."""
    features = _f(features)
    cv = build_cost_volume(features[:, ref_i], features[:, view_i], cams, depth_num, depth_start,
                           depth_interval, ref_id=0, view_id=view_i, inverse_depth=inverse_depth)
    prob, filt = cost_volume_reasoning(cv, weights, output_filtered_cost=True)
    return prob2depth(prob, depth_num, depth_start, depth_interval), prob, filt


def TVSNet_base_siamese(features, cams, depth_num, depth_start, depth_interval, view_i, weights,
                        ref_i=0, inverse_depth=True):
    """model.py:398-417 with features in place of images."""
    features = _f(features)
    depth_b2, prob, filt = TVSNet_base(features, cams, depth_num, depth_start, depth_interval,
                                       view_i, weights, ref_i, inverse_depth)
    cv_view = build_cost_volume(features[:, view_i], features[:, ref_i], cams, depth_num, depth_start,
                                depth_interval, ref_id=view_i, view_id=0, inverse_depth=inverse_depth)
    prob_view = cost_volume_reasoning(cv_view, weights)
    depth_view = prob2depth(prob_view, depth_num, depth_start, depth_interval)
    return depth_b2, prob, filt, depth_view


def run_multiview_stage12(features, cams, depth_num, weights, inverse_depth=True, siamese=True):
    """example.py:144-158: stage I per source view, host np.stack(axis=-1), stage II
    (AAM1 keepchannel -> output_conv -> prob2depth).  Returns dict of the stage outputs."""
    cams = _f(cams)
    B, N = cams.shape[:2]
    ds = cams[:, 0, 1, 3, 0]
    di = cams[:, 0, 1, 3, 1]
    filt, probs, dviews = [], [], []
    for view_i in range(1, N):
        if siamese:
            _, p, f, dv = TVSNet_base_siamese(features, cams, depth_num, ds, di, view_i, weights,
                                              inverse_depth=inverse_depth)
            dviews.append(dv)
        else:
            _, p, f = TVSNet_base(features, cams, depth_num, ds, di, view_i, weights,
                                  inverse_depth=inverse_depth)
        filt.append(f)
        probs.append(p)
    filt = np.stack(filt, axis=-1)
    cost_agg = cost_volume_aggregation(filt, weights, keepchannel=True)
    prob_agg = output_conv(cost_agg, weights)
    depth_init, depth_up = prob2depth_upsample(prob_agg, depth_num, ds, di)
    return dict(filtered_cost_volumes=filt, prob_volumes=np.stack(probs, axis=-1), depth_views=dviews,
                cost_volume_agg=cost_agg, prob_volume_agg=prob_agg, depth_agg_init=depth_init,
                depth_agg_init_up=depth_up)
