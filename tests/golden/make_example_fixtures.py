"""Copies the INPUT fixtures of BASELINE.json configs[0] (atvsnet/example.py on the bundled scenes, 3 views) out of the
read-only reference checkout, so that the -m gpu tests can run where /root/reference does not exist:
example/0/{0,1,2}.jpg + {0,1,2}_cam.npy (multi-view schedule with 2 sources, 960x640) and example/2/{0,1}.jpg + cams
(two-view network, 640x480).  Data files only (inputs of the reference's demo), no reference source code.

    python tests/golden/make_example_fixtures.py
"""
import os
import shutil

SRC = '/root/reference/example'
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'example')

for scene, views in (('0', 3), ('2', 2)):
    os.makedirs(os.path.join(DST, scene), exist_ok=True)
    for v in range(views):
        for name in ('%d.jpg' % v, '%d_cam.npy' % v):
            shutil.copyfile(os.path.join(SRC, scene, name), os.path.join(DST, scene, name))
            os.chmod(os.path.join(DST, scene, name), 0o644)
print(sorted(os.listdir(os.path.join(DST, '0'))), sorted(os.listdir(os.path.join(DST, '2'))))
