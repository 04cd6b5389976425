"""CPU: on-disk formats (a-tvsnet_b200/preprocess.py) - camera text, PFM, pair.txt - round trips
and hand-written files; the bundled example cameras (example/*/N_cam.npy) survive write_cam -> load_cam."""
import importlib.util
import io
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('atvs_preprocess', os.path.join(ROOT, 'a-tvsnet_b200', 'preprocess.py'))
P = importlib.util.module_from_spec(spec)
spec.loader.exec_module(P)

CAM_TXT = """extrinsic
1 0 0 0.5
0 1 0 -2
0 0 1 3.25
0 0 0 1

intrinsic
500.5 0 320
0 501.5 240
0 0 1

%s
"""


def test_load_cam_depth_field_variants():
    c = P.load_cam(io.StringIO(CAM_TXT % '425.0 2.5'), max_d=192)                 # 29 words: num = max_d
    assert c.shape == (2, 4, 4) and c[0, 0, 3] == 0.5 and c[0, 2, 3] == 3.25 and c[1, 1, 1] == 501.5
    assert list(c[1, 3]) == [425.0, 2.5, 192.0, 425.0 + 2.5 * 192]
    c = P.load_cam(io.StringIO(CAM_TXT % '425.0 2.5 64'), interval_scale=2)       # 30 words: end derived
    assert list(c[1, 3]) == [425.0, 5.0, 64.0, 425.0 + 5.0 * 64]
    c = P.load_cam(io.StringIO(CAM_TXT % '0.05 0.0033 128 0.4724'))               # 31 words
    assert list(c[1, 3]) == [0.05, 0.0033, 128.0, 0.4724]
    assert list(P.load_cam(io.StringIO(CAM_TXT % ''))[1, 3]) == [0, 0, 0, 0]
    assert np.all(c[1, :3, 3] == 0) and np.all(c[0, 3] == [0, 0, 0, 1])


def test_write_cam_round_trip(tmp_path):
    ref = '/root/reference/example/0/2_cam.npy'
    if os.path.exists(ref):
        cam = np.load(ref).astype(np.float64)
    else:
        rng = np.random.default_rng(0)
        cam = np.zeros((2, 4, 4))
        cam[0] = np.eye(4)
        cam[0, :3] = rng.standard_normal((3, 4))
        cam[1, :3, :3] = [[131.7, 0, 120], [0, 132.1, 80], [0, 0, 1]]
        cam[1, 3] = [0.05, 0.0033, 128, 0.4724]
    f = str(tmp_path / 'c_cam.txt')
    P.write_cam(f, cam)
    back = P.load_cam(f)
    assert np.array_equal(back[0], cam[0]) and np.array_equal(back[1, :3, :3], cam[1, :3, :3])
    assert np.array_equal(back[1, 3], cam[1, 3])
    s = P.scale_camera(cam, 0.25)
    assert s[1, 0, 0] == cam[1, 0, 0] * 0.25 and s[1, 1, 2] == cam[1, 1, 2] * 0.25 and s[1, 2, 2] == cam[1, 2, 2]
    assert np.array_equal(s[0], cam[0]) and s is not cam


def test_pfm_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(1)
    for shape in ((5, 7), (4, 6, 1), (3, 5, 3)):
        img = rng.standard_normal(shape).astype(np.float32)
        f = str(tmp_path / 'a.pfm')
        P.write_pfm(f, img)
        back = P.load_pfm(f)
        assert np.array_equal(back, img.reshape(back.shape))
    raw = open(f, 'rb').read()
    assert raw.startswith(b'PF\n5 3\n-1.000000\n')                                # colour, width height, little endian
    # rows are stored bottom-up: the first stored row is the LAST image row
    first_row = np.frombuffer(raw[len(b'PF\n5 3\n-1.000000\n'):][:5 * 3 * 4], '<f4').reshape(5, 3)
    assert np.array_equal(first_row, img[-1])
    # big-endian file written by hand
    g = np.arange(6, dtype='>f4').reshape(2, 3)
    open(f, 'wb').write(b'Pf\n3 2\n1.0\n' + g.tobytes())
    assert np.array_equal(P.load_pfm(f), np.float32([[3, 4, 5], [0, 1, 2]]))
    with pytest.raises(Exception):
        P.write_pfm(f, img.astype(np.float64))
    open(f, 'wb').write(b'P5\n3 2\n1.0\n')
    with pytest.raises(Exception):
        P.load_pfm(f)


def test_pair_txt(tmp_path):
    d = tmp_path / 'dense'
    d.mkdir()
    (d / 'pair.txt').write_text('2\n0\n3 1 0.9 2 0.8 5 0.1\n7\n1 0 0.5\n')
    lst = P.gen_pipeline_mvs_list(str(d), view_num=3)
    assert len(lst) == 2
    assert [os.path.basename(p) for p in lst[0]] == ['00000000.jpg', '00000000_cam.txt', '00000001.jpg', '00000001_cam.txt',
                                                      '00000002.jpg', '00000002_cam.txt']
    assert [os.path.basename(p) for p in lst[1]] == ['00000007.jpg', '00000007_cam.txt', '00000000.jpg', '00000000_cam.txt']
    assert lst[0][1].endswith(os.path.join('cams', '00000000_cam.txt'))


def test_center_image():
    rng = np.random.default_rng(2)
    img = (rng.uniform(0, 255, (20, 30, 3))).astype(np.uint8)
    c = P.center_image(img)
    assert c.dtype == np.float32 and np.allclose(c.mean(axis=(0, 1)), 0, atol=1e-5) and np.allclose(c.std(axis=(0, 1)), 1, atol=1e-4)
