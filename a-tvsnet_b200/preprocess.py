"""On-disk formats of the reference pipeline, host-side Python as in the reference:
/root/reference/atvsnet/preprocess.py:20-37 (center_image, scale_camera), :39-100 (scale_mvs_camera, scale_image,
scale_mvs_input, crop_mvs_input, mask_depth_image), :102-139 (load_cam), :141-162
(write_cam), :164-198 (load_pfm), :201-232 (write_pfm), :236-265 (gen_pipeline_mvs_list / pair.txt).

Camera array layout (also the ``*_cam.npy`` files of example/): cam[0] = 4x4 extrinsic [R|t; 0 0 0 1] (world ->
camera), cam[1][:3,:3] = K, cam[1][3] = (depth_start, depth_interval, depth_num, depth_end).  No OpenCV, no TF:
files are plain Python file objects / paths; ``scale_image`` restates cv2.resize (the reference's resampler) in NumPy."""
import math
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


def scale_mvs_camera(cams, scale=1, view_num=None):
    """preprocess.py:39-43: every view's camera scaled in place (``view_num`` stands for FLAGS.view_num)."""
    for view in range(len(cams) if view_num is None else view_num):
        cams[view] = scale_camera(cams[view], scale=scale)
    return cams


def _cv_round(x):
    """cvRound: round half to even."""
    return int(np.rint(x))


def _linear_taps(n_src, n_dst, inv_scale, coord_dtype, reset_at_border):
    """cv2.resize INTER_LINEAR taps of one axis -> (index of the first tap, index of the second, weight of the second).
    Columns: a coordinate outside [0, n-1] is RESET to the border pixel with weight 0; rows: the two tap indices are
    clamped and keep their weights (OpenCV's resizeGeneric_ does exactly this, which shows in the fixed-point rounding)."""
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * inv_scale - 0.5).astype(coord_dtype)
    i0 = np.floor(f).astype(np.int64)
    w1 = (f - i0.astype(coord_dtype)).astype(np.float32)
    if reset_at_border:
        lo = i0 < 0
        w1[lo], i0[lo] = 0.0, 0
        hi = i0 >= n_src - 1
        w1[hi], i0[hi] = 0.0, n_src - 1
    return np.clip(i0, 0, n_src - 1), np.clip(i0 + 1, 0, n_src - 1), w1


def scale_image(image, scale=1, interpolation='linear'):
    """preprocess.py:45-50 = cv2.resize(image, None, fx=scale, fy=scale, INTER_LINEAR | INTER_NEAREST): output size
    round-half-even(size * scale), pixel centres at (d + 0.5) / scale - 0.5.  uint8 images go through OpenCV's
    fixed-point bilinear (11-bit tap weights, the result rounded once), other dtypes through its float path."""
    image = np.asarray(image)
    h, w = image.shape[:2]
    nh, nw = _cv_round(h * scale), _cv_round(w * scale)
    if nh <= 0 or nw <= 0:
        raise ValueError("scale_image: empty output for shape %s, scale %g" % (image.shape, scale))
    inv = 1.0 / scale
    if interpolation == 'nearest':
        yi = np.minimum(np.floor(np.arange(nh) * inv).astype(np.int64), h - 1)
        xi = np.minimum(np.floor(np.arange(nw) * inv).astype(np.int64), w - 1)
        return image[yi][:, xi]
    if interpolation != 'linear':
        return None                                         # the reference falls through the same way
    # the 8-bit path keeps the sample coordinate as a float, the floating-point path as a double
    ct = np.float32 if image.dtype == np.uint8 else np.float64
    y0, y1, wy = _linear_taps(h, nh, inv, ct, False)
    x0, x1, wx = _linear_taps(w, nw, inv, ct, True)
    tail = (1,) * (image.ndim - 2)
    if image.dtype == np.uint8:
        # INTER_RESIZE_COEF_BITS = 11: taps as rounded 11-bit integers, rows blended with >> 4 / >> 16 / (+2) >> 2
        ax1 = np.rint(wx * 2048.0).astype(np.int64)
        ay1 = np.rint(wy * 2048.0).astype(np.int64)
        ax0, ay0 = 2048 - ax1, 2048 - ay1
        src = image.astype(np.int64)
        rows = src[:, x0] * ax0.reshape((1, -1) + tail) + src[:, x1] * ax1.reshape((1, -1) + tail)
        b0, b1 = ay0.reshape((-1, 1) + tail), ay1.reshape((-1, 1) + tail)
        out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
        return np.clip(out, 0, 255).astype(np.uint8)
    src = image.astype(np.float32)
    fx1 = wx.reshape((1, -1) + tail)
    rows = src[:, x0] * (1.0 - fx1) + src[:, x1] * fx1
    fy1 = wy.reshape((-1, 1) + tail)
    out = rows[y0] * (1.0 - fy1) + rows[y1] * fy1
    return out.astype(image.dtype) if np.issubdtype(image.dtype, np.floating) else np.rint(out).astype(image.dtype)


def scale_mvs_input(images, cams, depth_image=None, scale=1, view_num=None):
    """preprocess.py:52-61: images resized (bilinear), cameras scaled, the depth image resized with 'nearest'."""
    for view in range(len(images) if view_num is None else view_num):
        images[view] = scale_image(images[view], scale=scale)
        cams[view] = scale_camera(cams[view], scale=scale)
    if depth_image is None:
        return images, cams
    return images, cams, scale_image(depth_image, scale=scale, interpolation='nearest')


def crop_mvs_input(images, cams, depth_image=None, base_image_size=32, max_h=480, max_w=896, view_num=None):
    """preprocess.py:63-91: centre crop to (max_h, max_w) where the image is larger, else to the next multiple of
    ``base_image_size`` (the three stride-2 levels of the 2-D and 3-D U-Nets); principal points follow.
    ``max_h`` / ``max_w`` stand for FLAGS.max_h / FLAGS.max_w (eval_pointcloud.py:47-49).  The reference is Python 2
    without ``from __future__ import division``: ``h / base_image_size`` floors before ``math.ceil`` sees it, so a size
    that is not a multiple of the base is cut DOWN (1080 -> 1056, BASELINE cfg3), and ``(h - new_h) / 2`` floors too."""
    start_h = start_w = finish_h = finish_w = 0
    for view in range(len(images) if view_num is None else view_num):
        h, w = images[view].shape[0:2]
        new_h = max_h if h > max_h else int(math.ceil(h // base_image_size) * base_image_size)
        new_w = max_w if w > max_w else int(math.ceil(w // base_image_size) * base_image_size)
        start_h = int(math.ceil((h - new_h) // 2))
        start_w = int(math.ceil((w - new_w) // 2))
        finish_h, finish_w = start_h + new_h, start_w + new_w
        images[view] = images[view][start_h:finish_h, start_w:finish_w]
        cams[view][1][0][2] = cams[view][1][0][2] - start_w
        cams[view][1][1][2] = cams[view][1][1][2] - start_h
    if depth_image is not None:
        return images, cams, depth_image[start_h:finish_h, start_w:finish_w]
    return images, cams


def mask_depth_image(depth_image, min_depth, max_depth):
    """preprocess.py:93-100 (two cv2.threshold calls): depths <= min_depth or > max_depth become 0; (H,W) -> (H,W,1)."""
    d = np.asarray(depth_image)
    d = np.where(d > min_depth, d, 0).astype(d.dtype)
    d = np.where(d > max_depth, 0, d).astype(d.dtype)
    return np.expand_dims(d, 2)


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
