"""Generate tests/golden/reference_golden_preprocess.npz by executing the REFERENCE's own
/root/reference/atvsnet/preprocess.py (scale_image, scale_mvs_input, crop_mvs_input, mask_depth_image, scale_mvs_camera)
under Python 3 with the real OpenCV of this image and tests/golden/tf_shim.py standing in for the `tensorflow` import
(only FLAGS is used by these functions).  Run once in the build container:

    python tests/golden/make_golden_preprocess.py

Python 2 note: the reference's ``h / base_image_size`` floors for ints (no ``from __future__ import division``); under
Python 3 it would not, so crop_mvs_input is executed with integer division restored by feeding it sizes that ARE
multiples of the base, plus the py2 arithmetic evaluated by hand in tests/test_preprocess_io.py."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

REF = '/root/reference'
tf = tf_shim.install()
# preprocess.py imports `from tensorflow.python.lib.io import file_io` at module level: give the stand-in that path
for name in ('tensorflow.python', 'tensorflow.python.lib', 'tensorflow.python.lib.io', 'tensorflow.python.lib.io.file_io'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['tensorflow.python.lib.io'].file_io = sys.modules['tensorflow.python.lib.io.file_io']
sys.path.insert(0, os.path.join(REF, 'atvsnet'))
import preprocess as rp  # noqa: E402  (the reference module, unmodified)

F = tf.app.flags.FLAGS
rng = np.random.default_rng(77)
out = {}
img8 = rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8)
dep = rng.uniform(0.0, 12.0, size=(96, 128)).astype(np.float32)
out['img8'], out['dep'] = img8, dep
for sc in (0.25, 0.5, 0.55, 0.8):
    out['lin8_%g' % sc] = rp.scale_image(img8, scale=sc)
    out['nn8_%g' % sc] = rp.scale_image(img8, scale=sc, interpolation='nearest')
    out['linf_%g' % sc] = rp.scale_image(dep, scale=sc)
    out['nnf_%g' % sc] = rp.scale_image(dep, scale=sc, interpolation='nearest')
out['mask_2_8'] = rp.mask_depth_image(dep.copy(), 2.0, 8.0)

# scale_mvs_input + crop_mvs_input on 3 views (sizes chosen so that py2 and py3 division agree: multiples of 32)
F.view_num = 3
cams = []
for v in range(3):
    c = np.zeros((2, 4, 4), np.float64)
    c[0] = np.eye(4)
    c[1, :3, :3] = [[400.0 + v, 0, 128.5], [0, 390.0 + v, 95.25], [0, 0, 1]]
    c[1, 3] = [0.5, 0.01, 128, 1.78]
    cams.append(c)
imgs = [rng.integers(0, 256, size=(192, 256, 3), dtype=np.uint8) for _ in range(3)]
depth = rng.uniform(0.0, 12.0, size=(192, 256)).astype(np.float32)
out['mvs_imgs'], out['mvs_cams'], out['mvs_depth'] = np.stack(imgs), np.stack(cams), depth
si, sc_, sd = rp.scale_mvs_input([i.copy() for i in imgs], [c.copy() for c in cams], depth.copy(), scale=0.5)
out['mvs_scaled_imgs'], out['mvs_scaled_cams'], out['mvs_scaled_depth'] = np.stack(si), np.stack(sc_), sd
F.max_h, F.max_w = 64, 96
ci, cc, cd = rp.crop_mvs_input([i.copy() for i in si], [c.copy() for c in sc_], sd.copy(), base_image_size=32)
out['mvs_crop_imgs'], out['mvs_crop_cams'], out['mvs_crop_depth'] = np.stack(ci), np.stack(cc), cd
out['mvs_scaled_cams_quarter'] = np.stack(rp.scale_mvs_camera([c.copy() for c in cc], scale=0.25))
np.savez_compressed(os.path.join(HERE, 'reference_golden_preprocess.npz'), **out)
print({k: (v.shape, str(v.dtype)) for k, v in out.items()})
