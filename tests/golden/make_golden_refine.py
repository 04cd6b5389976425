"""Generate tests/golden/refine_variables.json and reference_golden_refine.npz by executing the REFERENCE's own refinement
stage (/root/reference/atvsnet/model.py:227-339, 428-441; homography_warping.py:275-387; cnn_wrapper/atvsnet.py:245-251,
295-336) under Python 3 on tests/golden/tf_shim.py:

    python tests/golden/make_golden_refine.py

Pass 1 records the variables the graph asks for, pass 2 runs on gen_common.named_weights('refine_variables.json', 5).
Build container only (reads /root/reference)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

REF = '/root/reference'
tf = tf_shim.install()
tf_shim.install_2d(tf)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'atvsnet'))
import model as rmodel  # noqa: E402
import homography_warping as rhw  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location('syn', os.path.join(ROOT, 'a-tvsnet_b200', 'synthetic.py'))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

T = tf_shim._t
F32 = np.float32
rng = np.random.default_rng(31)
B, N, H, W, D = 1, 3, 32, 64, 8
h, w = H // 4, W // 4
cams = syn.orbit_cams(N, h, w, D)[None]
ds, di = cams[:, 0, 1, 3, 0].copy(), cams[:, 0, 1, 3, 1].copy()
imgs = (127 + 50 * rng.standard_normal((B, N, H, W, 3))).clip(0, 255).astype(F32)
depth_b2 = (ds[0] + di[0] * rng.uniform(0.5, D - 1.5, (B, h, w, 1))).astype(F32)
depth_view = (ds[0] + di[0] * rng.uniform(0.5, D - 1.5, (B, h, w, 1))).astype(F32)
depth_view[0, 0, :3] = 0.0                      # invalid (zero) inverse depths exercise the 1e-10 masks
prob = rng.standard_normal((B, D, h, w)).astype(F32)
cost = rng.standard_normal((B, D, h, w, 8)).astype(F32)
view_i = 2


def run():
    return rmodel.TVSNet_refine(T(depth_b2), T(depth_view), T(prob), T(cost), T(imgs), T(cams), D, T(ds), T(di), view_i)


tf_shim.AUTO_VARS = {}
run()
shapes = {k: list(v) for k, v in tf_shim.AUTO_VARS.items()}
json.dump(shapes, open(os.path.join(HERE, 'refine_variables.json'), 'w'), indent=0, sort_keys=True)
tf_shim.AUTO_VARS = None

import gen_common  # noqa: E402
tf_shim.VARIABLES.clear()
tf_shim.VARIABLES.update(gen_common.named_weights('refine_variables.json', 5))
rp, rc = run()
out = dict(cams=cams, images=imgs, depth_b2=depth_b2, depth_view=depth_view, prob=prob, cost=cost, view_i=np.int32(view_i),
           refined_prob=np.asarray(rp), refined_cost=np.asarray(rc))
out['transform_depth'] = np.asarray(rhw.transform_depth(T(depth_view), T(cams[:, view_i]), T(cams[:, 0])))
out['visual_hull'] = np.asarray(rhw.get_visual_hull(T(np.stack([depth_b2, depth_view], 1)[..., 0]), T(cams), D, T(ds), T(di),
                                                    ref_id=0, view_num=2))
rf, vf = rmodel.extract_feature_shallow(T(imgs), 0, view_i)
out['shallow_ref'], out['shallow_view'] = np.asarray(rf), np.asarray(vf)
np.savez_compressed(os.path.join(HERE, 'reference_golden_refine.npz'), **out)
print('variables', len(shapes), 'wrote', len(out), 'arrays,', sum(np.asarray(v).nbytes for v in out.values()) / 1e6, 'MB raw')
