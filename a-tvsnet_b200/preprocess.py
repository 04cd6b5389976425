"""On-disk formats of the reference pipeline, host-side Python as in the reference:
/root/reference/atvsnet/preprocess.py:20-37 (center_image, scale_camera), :102-139 (load_cam), :141-162
(write_cam), :164-198 (load_pfm), :201-232 (write_pfm), :236-265 (gen_pipeline_mvs_list / pair.txt).

Camera array layout (also the ``*_cam.npy`` files of example/): cam[0] = 4x4 extrinsic [R|t; 0 0 0 1] (world ->
camera), cam[1][:3,:3] = K, cam[1][3] = (depth_start, depth_interval, depth_num, depth_end).  No OpenCV, no TF:
files are plain Python file objects / paths."""
import os
import re
import sys

import numpy as np


def center_image(img):
    """per-image, per-channel standardisation (preprocess.py:20-25)."""
    img = img.astype(np.float32)
    var = np.var(img, axis=(0, 1), keepdims=True)
    mean = np.mean(img, axis=(0, 1), keepdims=True)
    return (img - mean) / (np.sqrt(var) + 0.00000001)


def scale_camera(cam, scale=1):
    """focal lengths and principal point times ``scale`` (preprocess.py:27-37); returns a copy."""
    new_cam = np.copy(cam)
    for r, c in ((0, 0), (1, 1), (0, 2), (1, 2)):
        new_cam[1][r][c] = cam[1][r][c] * scale
    return new_cam


def load_cam(file, interval_scale=1, max_d=128):
    """MVSNet camera text: 'extrinsic' + 16 numbers, 'intrinsic' + 9 numbers, then 2, 3 or 4 depth fields
    (start, interval[, num[, end]]); missing fields are derived as in preprocess.py:118-137 (``max_d`` stands for
    FLAGS.max_d).  ``file``: open text file or path."""
    if isinstance(file, (str, bytes, os.PathLike)):
        with open(file) as f:
            return load_cam(f, interval_scale, max_d)
    words = file.read().split()
    cam = np.zeros((2, 4, 4))
    cam[0] = np.array(words[1:17], dtype=np.float64).reshape(4, 4)
    cam[1, :3, :3] = np.array(words[18:27], dtype=np.float64).reshape(3, 3)
    n = len(words)
    if n in (29, 30, 31):
        cam[1][3][0] = float(words[27])
        cam[1][3][1] = float(words[28]) * interval_scale
        cam[1][3][2] = max_d if n == 29 else float(words[29])
        cam[1][3][3] = float(words[30]) if n == 31 else cam[1][3][0] + cam[1][3][1] * cam[1][3][2]
    return cam


def write_cam(file, cam):
    """inverse of load_cam with all four depth fields (preprocess.py:141-162)."""
    with open(file, 'w') as f:
        f.write('extrinsic\n')
        for i in range(4):
            f.write(''.join(str(cam[0][i][j]) + ' ' for j in range(4)) + '\n')
        f.write('\nintrinsic\n')
        for i in range(3):
            f.write(''.join(str(cam[1][i][j]) + ' ' for j in range(3)) + '\n')
        f.write('\n' + ' '.join(str(cam[1][3][j]) for j in range(4)) + '\n')


def load_pfm(file):
    """PFM ('Pf' grey / 'PF' colour; negative scale = little endian; rows stored bottom-up) -> float32 array, top row
    first (preprocess.py:164-198).  ``file``: open binary file or path."""
    if isinstance(file, (str, os.PathLike)):
        with open(file, 'rb') as f:
            return load_pfm(f)
    header = file.readline().decode('latin-1').rstrip()
    if header not in ('PF', 'Pf'):
        raise Exception('Not a PFM file.')
    m = re.match(r'^(\d+)\s(\d+)\s$', file.readline().decode('latin-1'))
    if not m:
        raise Exception('Malformed PFM header.')
    width, height = int(m.group(1)), int(m.group(2))
    scale = float(file.readline().decode('latin-1').rstrip())
    data = np.frombuffer(file.read(), '<f4' if scale < 0 else '>f4')
    shape = (height, width, 3) if header == 'PF' else (height, width)
    return np.ascontiguousarray(np.flipud(data.reshape(shape)).astype(np.float32))


def write_pfm(file, image, scale=1):
    """float32 (H,W) | (H,W,1) | (H,W,3) -> PFM, native byte order (preprocess.py:201-232)."""
    if image.dtype.name != 'float32':
        raise Exception('Image dtype must be float32.')
    if len(image.shape) == 3 and image.shape[2] == 3:
        color = True
    elif len(image.shape) == 2 or (len(image.shape) == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception('Image must have H x W x 3, H x W x 1 or H x W dimensions.')
    image = np.flipud(image)
    endian = image.dtype.byteorder
    if endian == '<' or (endian == '=' and sys.byteorder == 'little'):
        scale = -scale
    with open(file, 'wb') as f:
        f.write(('PF\n' if color else 'Pf\n').encode())
        f.write(('%d %d\n' % (image.shape[1], image.shape[0])).encode())
        f.write(('%f\n' % scale).encode())
        f.write(np.ascontiguousarray(image).tobytes())


def gen_pipeline_mvs_list(dense_folder, view_num=5):
    """pair.txt -> per reference image [ref_image, ref_cam, view_image, view_cam, ...] with at most ``view_num`` - 1
    source views (preprocess.py:236-265; ``view_num`` stands for FLAGS.view_num)."""
    image_folder = os.path.join(dense_folder, 'images')
    cam_folder = os.path.join(dense_folder, 'cams')
    tokens = open(os.path.join(dense_folder, 'pair.txt')).read().split()
    mvs_list = []
    pos = 1
    for _ in range(int(tokens[0])):
        ref_index = int(tokens[pos])
        n_all = int(tokens[pos + 1])
        pos += 2
        paths = [os.path.join(image_folder, '%08d.jpg' % ref_index), os.path.join(cam_folder, '%08d_cam.txt' % ref_index)]
        for v in range(min(view_num - 1, n_all)):
            idx = int(tokens[pos + 2 * v])
            paths += [os.path.join(image_folder, '%08d.jpg' % idx), os.path.join(cam_folder, '%08d_cam.txt' % idx)]
        pos += 2 * n_all
        mvs_list.append(paths)
    return mvs_list
