"""NumPy / torch-CPU fp32 restatement of the refinement stage (stages III of example.py:160-172):
/root/reference/atvsnet/homography_warping.py:275-326 (transform_depth), :329-387 (get_visual_hull),
/root/reference/atvsnet/model.py:143-154 (extract_feature_shallow), :227-339 (refinement), :428-441 (TVSNet_refine),
/root/reference/cnn_wrapper/atvsnet.py:245-251 (ResNetDS2SPP_shallow_f16), :295-336 (CostVolRefineNet).

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for SURVEY.md section 8(f) row N2, written ahead of any CUDA
code for that row.  Pinned against the reference's own functions executed on tests/golden/tf_shim.py
(tests/golden/make_golden_refine.py -> reference_golden_refine.npz, refine_variables.json).

Reference behaviours kept on purpose:
  * the per-plane geometric error of the view is multiplied by the validity mask TILED to the 16 feature channels,
    so `geo_group` has 1 + 16 + 1 + 1 = 19 channels (model.py:297-300, the kernel is [3,3,3,19,8]);
  * get_visual_hull walks `range(view_num)` over the CAMERA list with view_num = num_depths = 2, i.e. it always pairs
    the transformed depth of slot 1 with cams[:, 1], whatever `view_id` is (homography_warping.py:344-357);
  * nearest-neighbour warps of invalid pixels read element [0,0] (no zeroing), homography_warping.py:45-56.
"""
import numpy as np

from . import fem
from . import homography_warping as hw
from . import network as net

F32 = np.float32


def _f(x):
    return np.asarray(x, dtype=F32)


def transform_depth(left_depth, left_cam, right_cam, inverse_depth=True):
    """homography_warping.py:275-326: depth values of the left view expressed in the right camera's frame, kept on the
    LEFT pixel grid.  left_depth (B,H,W,1) (inverse depth when ``inverse_depth``)."""
    d = _f(left_depth)
    shape = d.shape
    B, H, W = shape[0], shape[1], shape[2]
    left_cam, right_cam = _f(left_cam), _f(right_cam)
    R_l, R_r = left_cam[:, 0, :3, :3], right_cam[:, 0, :3, :3]
    t_l, t_r = left_cam[:, 0, :3, 3:4], right_cam[:, 0, :3, 3:4]
    K_l, K_r = left_cam[:, 1, :3, :3], right_cam[:, 1, :3, :3]
    R_l_T = np.transpose(R_l, (0, 2, 1))
    c_l = -np.matmul(R_l_T, t_l)
    grid = hw.get_pixel_grids(H, W).reshape(1, 3, -1)
    if inverse_depth:
        valid = d > F32(1e-10)
        d = np.clip(d, F32(1e-10), d.max())
        d = F32(1.0) / d
        d = d * valid.astype(F32)
    pts = grid * d.reshape(B, 1, H * W)
    mat = np.matmul(K_r, np.matmul(R_r, np.matmul(R_l_T, np.linalg.inv(K_l).astype(F32))))
    vec = np.matmul(K_r, np.matmul(R_r, c_l)) + np.matmul(K_r, t_r)
    xyz = np.matmul(mat, pts) + vec
    out = xyz[:, 2, :].reshape(shape)
    if inverse_depth:
        out = np.clip(out, F32(1e-10), out.max())
        out = F32(1.0) / out
        out = out * valid.astype(F32)
    return out.astype(F32)


def get_visual_hull(depth_images, cams, depth_num, depth_start, depth_interval, ref_id=0, view_num=2,
                    inverse_depth=True):
    """homography_warping.py:329-387: depth_images (B,N,H,W), cams (B,*,2,4,4) -> (B,D,H,W,1): per plane the fraction
    of views whose (transformed, nearest-warped) depth lies behind the plane."""
    depth_images, cams = _f(depth_images), _f(cams)
    ds, di = _f(depth_start), _f(depth_interval)
    order = list(range(view_num))
    order[0], order[ref_id] = ref_id, 0
    ref_cam = cams[:, ref_id]
    ref_depth = depth_images[:, ref_id]
    homos, trans = [], []
    for v in order[1:]:
        view_cam = cams[:, v]
        homos.append(hw.get_homographies(ref_cam, view_cam, depth_num, ds, di, inverse_depth=inverse_depth))
        trans.append(transform_depth(depth_images[:, v][..., None], view_cam, ref_cam, inverse_depth)[..., 0])
    planes = []
    for d in range(depth_num):
        cur = (ds + di * F32(d)).reshape(-1, 1, 1)
        sl = np.ones_like(ref_depth) * cur
        hull = (ref_depth > 0).astype(F32) * ((ref_depth > sl) if inverse_depth else (sl > ref_depth)).astype(F32)
        for i in range(view_num - 1):
            wd = hw.homography_warping(trans[i][..., None], homos[i][:, d], method='nearest')[..., 0]
            hull = hull + (wd > 0).astype(F32) * ((wd > sl) if inverse_depth else (sl > wd)).astype(F32)
        planes.append(hull)
    return (np.stack(planes, axis=1) / F32(view_num))[..., None].astype(F32)


def shallow_features(image, w):
    """cnn_wrapper/atvsnet.py:245-251: res_block(3, 16, 3 blocks, stride 4) + 1x1 conv (linear, no bias)."""
    x = fem.res_block(_f(image), w, 'global_refine_conv0_x', 16, 3, 4, 1)
    return fem.conv2d(x, w['global_refine_shallow_feature/kernel'], 1, 1)


def _conv_bn(x, w, name, stride=1):
    return net.relu(net.batch_norm_train(net.conv3d(x, w[name + '/conv3d/kernel'], stride)))


def _deconv_bn(x, w, name):
    return net.relu(net.batch_norm_train(net.deconv3d(x, w[name + '/conv3d_transpose/kernel'], 2)))


def CostVolRefineNet(photo_group, geo_group, prob_vol, vis_hull, w):
    """cnn_wrapper/atvsnet.py:295-336 -> (global_refine_3dconv6_1 (B,D,H,W,8), global_refined_cost_vol (B,D,H,W,1))."""
    p = 'global_refine_'
    cat = np.concatenate([_conv_bn(photo_group, w, p + 'photo_3dconv'), _conv_bn(geo_group, w, p + 'geo_3dconv'),
                          _conv_bn(prob_vol, w, p + 'prob_3dconv'), _conv_bn(vis_hull, w, p + 'vishull_3dconv')], axis=-1)
    c10 = _conv_bn(cat, w, p + '3dconv1_0', 2)
    c20 = _conv_bn(c10, w, p + '3dconv2_0', 2)
    c30 = _conv_bn(c20, w, p + '3dconv3_0', 2)
    c01 = _conv_bn(cat, w, p + '3dconv0_1')
    c11 = _conv_bn(c10, w, p + '3dconv1_1')
    c21 = _conv_bn(c20, w, p + '3dconv2_1')
    c31 = _conv_bn(c30, w, p + '3dconv3_1')
    c41 = _deconv_bn(c31, w, p + '3dconv4_0') + c21
    c51 = _deconv_bn(c41, w, p + '3dconv5_0') + c11
    c61 = _deconv_bn(c51, w, p + '3dconv6_0') + c01
    return c61, net.conv3d(c61, w['global_refined_cost_vol/kernel'], 1)


def refinement(init_depth_images, cams, depth_num, depth_start, depth_interval, images, prob_vol, ref_id, view_id, w,
               num_depths=2, depth_ref_id=None, depth_view_id=None, inverse_depth=True, return_groups=False):
    """model.py:227-339.  init_depth_images (B,2,h,w,1), images (B,N,H,W,3), prob_vol (B,D,h,w) ->
    (cost residual (B,D,h,w,8), prob residual (B,D,h,w))."""
    depth_ref_id = ref_id if depth_ref_id is None else depth_ref_id
    depth_view_id = view_id if depth_view_id is None else depth_view_id
    init_depth_images, cams, images = _f(init_depth_images), _f(cams), _f(images)
    ds, di = _f(depth_start), _f(depth_interval)
    D = int(depth_num)
    d_ref = init_depth_images[:, depth_ref_id]
    d_view = init_depth_images[:, depth_view_id]
    ref_cam, view_cam = cams[:, ref_id], cams[:, view_id]
    d_view_t = transform_depth(d_view, view_cam, ref_cam, inverse_depth)
    H_ = hw.get_homographies(ref_cam, view_cam, D, ds, di, inverse_depth=inverse_depth)
    ref_f = shallow_features(images[:, ref_id], w)
    view_f = shallow_features(images[:, view_id], w)
    C = ref_f.shape[-1]
    dsb, dib = ds.reshape(-1, 1, 1, 1), di.reshape(-1, 1, 1, 1)
    photo, geo_ref, geo_view = [], [], []
    for d in range(D):
        wv, m = hw.homography_warping(view_f, H_[:, d], output_mask=True)
        photo.append(np.abs(wv - ref_f) * np.tile(m, (1, 1, 1, C)).astype(F32))
        val = dsb + F32(d) * dib
        geo_ref.append(np.abs(d_ref - val) / dib / F32(D))
        wd, mv = hw.homography_warping(d_view_t, H_[:, d], output_mask=True)
        geo_view.append((np.abs(wd - val) / dib / F32(D)) * np.tile(mv, (1, 1, 1, C)).astype(F32))
    cost_photo = np.stack(photo, axis=1)
    cost_geo = np.concatenate([np.stack(geo_ref, axis=1), np.stack(geo_view, axis=1)], axis=-1)
    wf, mp = hw.homography_warping_by_depth(view_f, ref_cam, view_cam, d_ref, output_mask=True, inverse_depth=inverse_depth)
    photo_err = np.abs(wf - ref_f) * np.tile(mp, (1, 1, 1, C)).astype(F32)
    wg, mg = hw.homography_warping_by_depth(d_view_t, ref_cam, view_cam, d_ref, output_mask=True, method='nearest',
                                            inverse_depth=inverse_depth)
    geo_err = np.abs(wg - d_ref) * mg.astype(F32)
    tile = lambda t: np.tile(t[:, None], (1, D, 1, 1, 1))
    vis = get_visual_hull(init_depth_images[..., 0], cams, D, ds, di, ref_id=ref_id, view_num=num_depths,
                          inverse_depth=inverse_depth)
    photo_group = np.concatenate([cost_photo, tile(photo_err), tile(ref_f)], axis=-1)
    geo_group = np.concatenate([cost_geo, tile(geo_err), tile(d_ref)], axis=-1)
    c61, res = CostVolRefineNet(photo_group, geo_group, _f(prob_vol)[..., None], vis, w)
    if return_groups:
        return c61, res[..., 0], dict(photo_group=photo_group, geo_group=geo_group, vis_hull=vis, depth_view_trans=d_view_t)
    return c61, res[..., 0]


def TVSNet_refine(depth_b2, depth_view, prob_vol_b2, filtered_cost_volume, images, cams, depth_num, depth_start,
                  depth_interval, view_i, w, ref_i=0, inverse_depth=True):
    """model.py:428-441 -> (refined_prob_vol (B,D,h,w), refined_cost_volume (B,D,h,w,8))."""
    init = np.stack([_f(depth_b2), _f(depth_view)], axis=1)
    cost_res, prob_res = refinement(init, cams, depth_num, depth_start, depth_interval, images, prob_vol_b2, ref_i, view_i, w,
                                    num_depths=2, depth_ref_id=0, depth_view_id=1, inverse_depth=inverse_depth)
    return _f(prob_vol_b2) + prob_res, _f(filtered_cost_volume) + cost_res
