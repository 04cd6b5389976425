"""Generate tests/golden/reference_golden.npz by executing the REFERENCE's own source files
(/root/reference/atvsnet/{homography_warping,model}.py, /root/reference/cnn_wrapper/*.py)
under Python 3 on top of tests/golden/tf_shim.py.  Run once in the build container:

    python tests/golden/make_golden.py

The GPU box never runs this (it has no /root/reference); tests read only the .npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402
from gen_common import golden_weights, smooth  # noqa: E402

REF = '/root/reference'
tf = tf_shim.install()
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'atvsnet'))
import homography_warping as rhw  # noqa: E402  (the reference module, unmodified)
import model as rmodel  # noqa: E402

T = tf_shim._t
F32 = np.float32
out = {}
rng = np.random.default_rng(2024)

# ---- get_homographies on the bundled example cameras (example/0, views 0..4)
cams = np.stack([np.load(os.path.join(REF, 'example/0/%d_cam.npy' % i)) for i in range(5)]).astype(F32)
out['ex0_cams'] = cams
ds, di = cams[0, 1, 3, 0:1].copy(), cams[0, 1, 3, 1:2].copy()
for v in (1, 4):
    out['ex0_H_0_%d' % v] = np.asarray(rhw.get_homographies(T(cams[0:1]), T(cams[v:v + 1]), 128, T(ds), T(di)))
out['ex0_H_3_0'] = np.asarray(rhw.get_homographies(T(cams[3:4]), T(cams[0:1]), 16, T(ds), T(di)))
tf.app.flags.FLAGS.inverse_depth = False
out['ex0_H_0_2_depth'] = np.asarray(rhw.get_homographies(T(cams[0:1]), T(cams[2:3]), 8, T(F32([2.0])), T(F32([0.5]))))
tf.app.flags.FLAGS.inverse_depth = True

# ---- homography_warping / by_depth on a 40x60 feature map with the example cameras
# (K of the example is for 160x240 features: rescale to 40x60)
h, w, C = 40, 60, 8
scams = cams.copy()
scams[:, 1, 0, :3] *= w / 240.0
scams[:, 1, 1, :3] *= h / 160.0
out['warp_cams'] = scams
img = smooth(rng, (1, h, w, C))
out['warp_img'] = img
Hs = np.asarray(rhw.get_homographies(T(scams[0:1]), T(scams[1:2]), 8, T(ds), T(di * 16)))
out['warp_H'] = Hs
for d in (0, 3, 7):
    o, m = rhw.homography_warping(T(img), T(Hs[:, d]), output_mask=True)
    out['warp_bilinear_%d' % d], out['warp_mask_%d' % d] = np.asarray(o), np.asarray(m)
o, m = rhw.homography_warping(T(img[..., :1]), T(Hs[:, 3]), method='nearest', output_mask=True)
out['warp_nearest_3'], out['warp_nearest_mask_3'] = np.asarray(o), np.asarray(m)
inv_depth = (ds[0] + di[0] * 16 * (3.0 + 2.0 * smooth(rng, (1, h, w, 1)))).astype(F32)
out['bydepth_depth'] = inv_depth
o, m = rhw.homography_warping_by_depth(T(img), T(scams[0:1]), T(scams[1:2]), T(inv_depth), output_mask=True)
out['bydepth_out'], out['bydepth_mask'] = np.asarray(o), np.asarray(m)

# ---- build_cost_volume (model.py:157) incl. warp_ref
ref, view = smooth(rng, (1, h, w, C)), img
out['cv_ref'] = ref
cams_b = scams[None, :2]
out['cv_concat'] = np.asarray(rmodel.build_cost_volume(T(ref), T(view), T(cams_b), 4, T(ds), T(di * 32), 0, 1))
out['cv_concat_rev'] = np.asarray(rmodel.build_cost_volume(T(view), T(ref), T(cams_b), 4, T(ds), T(di * 32), 1, 0))
out['cv_warpref'] = np.asarray(rmodel.build_cost_volume(T(ref), T(view), T(cams_b), 2, T(ds), T(di * 32), 0, 1,
                                                        warp_ref=True))

# ---- CRM: cost_volume_reasoning (model.py:204) = StackedUNet_prob graph of the reference
tf_shim.VARIABLES.update(golden_weights(7))
x = smooth(rng, (1, 8, 16, 16, 64))
out['crm_in'] = x
p, f = rmodel.cost_volume_reasoning(T(x), output_filtered_cost=True)
out['crm_prob'], out['crm_filtered'] = np.asarray(p), np.asarray(f)
tower = rmodel.StackedUNet_prob({'data': T(x)}, is_training=True, reuse=tf.AUTO_REUSE)
for nm in ('conv_b0_1_0', 'conv_b0_0_1', 'conv_b0_3_1', 'conv_b0_4_0', 'conv_b0_6_0', 'conv_b1_0_0', 'conv_b1_5_0'):
    out['crm_' + nm] = np.asarray(tower.get_output_by_name(nm))
out['crm_filtered_only'] = np.asarray(rmodel.cost_volume_reasoning(T(x), output_prob=False))

# ---- AAM1 / AAM2 / output convs
xs = smooth(rng, (1, 8, 8, 8, 8, 3))
out['aam_in'] = xs
out['aam1_keep'] = np.asarray(rmodel.cost_volume_aggregation(T(xs), keepchannel=True))
out['aam1_prob'] = np.asarray(rmodel.cost_volume_aggregation(T(xs), keepchannel=False))
out['aam2_keep'] = np.asarray(rmodel.cost_volume_aggregation_refine(T(xs), keepchannel=True))
out['outconv'] = np.asarray(rmodel.output_conv(T(out['aam1_keep'])))
out['outconv_refine'] = np.asarray(rmodel.output_conv_refine(T(out['aam1_keep'])))

# ---- prob2depth / prob map / x4 upsample
vol = (3.0 * smooth(rng, (1, 16, 10, 12, 1))[..., 0]).astype(F32)
out['p2d_vol'] = vol
e, pm = rmodel.prob2depth(T(vol), 16, T(ds), T(di * 8), out_prob_map=True)
out['p2d_est'], out['p2d_prob'] = np.asarray(e), np.asarray(pm)
e, eu, pm, pmu = rmodel.prob2depth_upsample(T(vol), 16, T(ds), T(di * 8), out_prob_map=True)
out['p2d_est_up'], out['p2d_prob_up'] = np.asarray(eu), np.asarray(pmu)
out['p2d_start'], out['p2d_interval'] = ds, di * 8
out['ds'], out['di'] = ds, di

np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
print('wrote', len(out), 'arrays,', sum(v.nbytes for v in out.values()) / 1e6, 'MB raw')
