"""Generate tests/golden/fem_variables.json and reference_golden_fem.npz by executing the REFERENCE's own FEM graph
(/root/reference/cnn_wrapper/atvsnet.py:254-292 on network.py's layer set) under Python 3 on tests/golden/tf_shim.py:

    python tests/golden/make_golden_fem.py

Pass 1 lets the shim create every variable the graph asks for and records name -> shape (fem_variables.json: the
checkpoint variable list of SURVEY.md Appendix B, as the reference code itself spells it).  Pass 2 re-runs the graph
on gen_common.fem_weights(7) and stores the output plus intermediate layers.  Build container only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

REF = '/root/reference'
tf = tf_shim.install()
tf_shim.install_2d(tf)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, 'atvsnet'))
import model as rmodel  # noqa: E402

T = tf_shim._t
F32 = np.float32
rng = np.random.default_rng(77)
# a 96x128 "image" in the 0..255 range the reference feeds (example.py:332-336): features 24x32, SPP pools 64/32/16/8
img = (127.5 + 60.0 * rng.standard_normal((1, 96, 128, 3))).clip(0, 255).astype(F32)

tf_shim.AUTO_VARS = {}
rmodel.ResNetDS2SPP({'data': T(img)}, is_training=True, reuse=tf.AUTO_REUSE).get_output()
shapes = {k: list(v) for k, v in tf_shim.AUTO_VARS.items()}
json.dump(shapes, open(os.path.join(HERE, 'fem_variables.json'), 'w'), indent=0, sort_keys=True)
tf_shim.AUTO_VARS = None

import gen_common  # noqa: E402  (reads fem_variables.json)
tf_shim.VARIABLES.clear()
tf_shim.VARIABLES.update(gen_common.fem_weights(7))
net = rmodel.ResNetDS2SPP({'data': T(img)}, is_training=True, reuse=tf.AUTO_REUSE)
out = {'image': img, 'feature': np.asarray(net.get_output())}
for nm in ('conv0_2', 'conv0_x', 'conv1_x', 'conv2_x', 'conv3_x', 'branch_0', 'branch_3', 'fusion0'):
    out[nm] = np.asarray(net.get_output_by_name(nm))
np.savez_compressed(os.path.join(HERE, 'reference_golden_fem.npz'), **out)
print('variables', len(shapes), 'params', sum(int(np.prod(s)) for s in shapes.values()))
print('wrote', len(out), 'arrays,', sum(v.nbytes for v in out.values()) / 1e6, 'MB raw')
