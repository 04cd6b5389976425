"""NumPy fp32 restatement of /root/reference/atvsnet/homography_warping.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Function names, argument order and
tensor layouts are the reference's.  All arithmetic is IEEE fp32 with an explicit,
element-wise operation order (no BLAS, no FMA contraction) so that the CUDA kernels can
mirror it bit for bit:

    dot3(a, b)      = (a0*b0 + a1*b1) + a2*b2
    mm3(A, B)[i,j]  = dot3(A[i,:], B[:,j])
    inv3(K)         = cofactor(K)^T / det(K)    (each entry one IEEE division)

``inverse_depth`` replaces the reference's module-level ``FLAGS.inverse_depth``
(homography_warping.py:6,149,215); default True like example.py:47.
"""
import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- helpers
def _f(x):
    return np.asarray(x, dtype=F32)


def mm3(a, b):
    """(...,3,K) x (...,K,3) for K == 3 with the fixed fp32 order documented above."""
    a = _f(a)
    b = _f(b)
    out_shape = np.broadcast_shapes(a.shape[:-2], b.shape[:-2]) + (a.shape[-2], b.shape[-1])
    out = np.empty(out_shape, dtype=F32)
    for i in range(a.shape[-2]):
        for j in range(b.shape[-1]):
            out[..., i, j] = (a[..., i, 0] * b[..., 0, j] + a[..., i, 1] * b[..., 1, j]) \
                + a[..., i, 2] * b[..., 2, j]
    return out


def inv3(k):
    """3x3 inverse by cofactors in fp32 (stands in for tf.matrix_inverse,
    homography_warping.py:123,199,290)."""
    k = _f(k)
    a, b, c = k[..., 0, 0], k[..., 0, 1], k[..., 0, 2]
    d, e, f = k[..., 1, 0], k[..., 1, 1], k[..., 1, 2]
    g, h, i = k[..., 2, 0], k[..., 2, 1], k[..., 2, 2]
    c00 = e * i - f * h
    c01 = d * i - f * g
    c02 = d * h - e * g
    det = (a * c00 - b * c01) + c * c02
    out = np.empty_like(k)
    out[..., 0, 0] = c00 / det
    out[..., 0, 1] = (c * h - b * i) / det
    out[..., 0, 2] = (b * f - c * e) / det
    out[..., 1, 0] = (f * g - d * i) / det
    out[..., 1, 1] = (a * i - c * g) / det
    out[..., 1, 2] = (c * d - a * f) / det
    out[..., 2, 0] = c02 / det
    out[..., 2, 1] = (b * g - a * h) / det
    out[..., 2, 2] = (a * e - b * d) / det
    return out


# ------------------------------------------------------------------- reference API
def get_pixel_grids(height, width):
    """homography_warping.py:8-17.  (x+0.5, y+0.5, 1), x fastest, stacked as 3 rows
    flattened to (3*H*W,)."""
    x = np.arange(width, dtype=F32) + F32(0.5)
    y = np.arange(height, dtype=F32) + F32(0.5)
    xc, yc = np.meshgrid(x, y)
    xc = xc.reshape(-1)
    yc = yc.reshape(-1)
    return np.concatenate([xc, yc, np.ones_like(xc)], 0)


def _tf_round(x):
    # tf.round is round-half-to-even (np.rint as well)
    return np.rint(x)


def interpolate(image, x, y, output_mask=False, method='bilinear'):
    """homography_warping.py:31-104.  image (B,H,W,C); x,y (B*H*W,) in texture
    coordinates.  Returns (B*H*W, C) [+ bool mask (B*H*W,)]."""
    image = _f(image)
    B, H, W, C = image.shape
    x = _f(x) - F32(0.5)
    y = _f(y) - F32(0.5)
    with np.errstate(invalid='ignore'):
        valid = (x >= 0) & (y >= 0) & (x < F32(W - 1)) & (y < F32(H - 1))
        valid &= ~np.isnan(x) & ~np.isnan(y)
    vi = valid.astype(np.int32)
    vf = valid.astype(F32)
    b = np.repeat(np.arange(B, dtype=np.int32), H * W)

    def _to_i32(v):
        with np.errstate(invalid='ignore'):
            v = np.where(np.isfinite(v), v, F32(0))
            return np.clip(v, -2147483648.0, 2147483520.0).astype(np.int32)

    if method == 'nearest':
        x0 = _to_i32(_tf_round(x)) * vi
        y0 = _to_i32(_tf_round(y)) * vi
        out = image[b, y0, x0]
        return (out, valid) if output_mask else out

    x0 = _to_i32(np.floor(x))
    y0 = _to_i32(np.floor(y))
    x1 = x0 + 1
    y1 = y0 + 1
    with np.errstate(invalid='ignore'):
        x = x * vf
        y = y * vf
    x0 = np.clip(x0 * vi, 0, W - 1)
    x1 = np.clip(x1 * vi, 0, W - 1)
    y0 = np.clip(y0 * vi, 0, H - 1)
    y1 = np.clip(y1 * vi, 0, H - 1)
    pa = image[b, y0, x0]
    pb = image[b, y0, x1]
    pc = image[b, y1, x0]
    pd = image[b, y1, x1]
    x0f, x1f, y0f, y1f = (v.astype(F32) for v in (x0, x1, y0, y1))
    with np.errstate(invalid='ignore'):
        area_a = ((y1f - y) * (x1f - x))[:, None]
        area_b = ((y1f - y) * (x - x0f))[:, None]
        area_c = ((y - y0f) * (x1f - x))[:, None]
        area_d = ((y - y0f) * (x - x0f))[:, None]
        out = ((area_a * pa + area_b * pb) + area_c * pc) + area_d * pd
    return (out, valid) if output_mask else out


def get_homographies(left_cam, right_cam, depth_num, depth_start, depth_interval,
                     inverse_depth=True):
    """homography_warping.py:179-227.  cams (B,2,4,4); start/interval (B,) -> (B,D,3,3)."""
    left_cam = _f(left_cam)
    right_cam = _f(right_cam)
    depth_start = _f(depth_start).reshape(-1)
    depth_interval = _f(depth_interval).reshape(-1)
    R_l, R_r = left_cam[:, 0, :3, :3], right_cam[:, 0, :3, :3]
    t_l, t_r = left_cam[:, 0, :3, 3:4], right_cam[:, 0, :3, 3:4]
    K_l, K_r = left_cam[:, 1, :3, :3], right_cam[:, 1, :3, :3]
    B = R_l.shape[0]
    depth = depth_start[:, None] + np.arange(depth_num, dtype=F32)[None, :] * depth_interval[:, None]
    K_l_inv = inv3(K_l)
    R_l_T = np.transpose(R_l, (0, 2, 1))
    R_r_T = np.transpose(R_r, (0, 2, 1))
    fronto = R_l[:, 2:3, :]                                     # (B,1,3)

    def mv3(m, v):  # (B,3,3) x (B,3,1)
        return ((m[:, :, 0:1] * v[:, 0:1, :] + m[:, :, 1:2] * v[:, 1:2, :]) + m[:, :, 2:3] * v[:, 2:3, :])

    c_l = -mv3(R_l_T, t_l)
    c_r = -mv3(R_r_T, t_r)
    c_rel = c_r - c_l                                           # (B,3,1)
    temp = c_rel * fronto                                       # (B,3,3) outer product
    temp = temp[:, None]                                        # (B,1,3,3)
    dm = depth.reshape(B, depth_num, 1, 1)
    eye = np.eye(3, dtype=F32)[None, None]
    if inverse_depth:
        m0 = eye - temp * dm
    else:
        m0 = eye - temp / dm
    m1 = mm3(R_l_T, K_l_inv)[:, None]                           # (B,1,3,3)
    m2 = mm3(m0, m1)
    return mm3(K_r[:, None], mm3(R_r[:, None], m2))


def _warp_coords(homography, height, width):
    """homography_warping.py:237-257: texture coordinates of every pixel under H."""
    hmg = _f(homography)
    B = hmg.shape[0]
    grid = get_pixel_grids(height, width).reshape(3, -1)
    px, py = grid[0][None], grid[1][None]                       # (1,HW)
    h = hmg.reshape(B, 9, 1)
    xa = (h[:, 0] * px + h[:, 1] * py) + h[:, 2]
    ya = (h[:, 3] * px + h[:, 4] * py) + h[:, 5]
    z = (h[:, 6] * px + h[:, 7] * py) + h[:, 8]
    z = z + (z == 0).astype(F32) * F32(1e-7)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        return (xa / z).reshape(-1), (ya / z).reshape(-1)


def homography_warping(input_image, homography, method='bilinear', output_mask=False):
    """homography_warping.py:230-271.  (B,H,W,C), (B,3,3) -> (B,H,W,C) [, (B,H,W,1) bool]."""
    input_image = _f(input_image)
    B, H, W, C = input_image.shape
    xw, yw = _warp_coords(homography, H, W)
    if output_mask:
        out, mask = interpolate(input_image, xw, yw, output_mask=True, method=method)
        return out.reshape(B, H, W, C), mask.reshape(B, H, W, 1)
    return interpolate(input_image, xw, yw, method=method).reshape(B, H, W, C)


def homography_warping_by_depth(input_image, left_cam, right_cam, depth_image,
                                output_mask=False, method='bilinear', inverse_depth=True):
    """homography_warping.py:108-176.  depth_image (B,H,W,1) holds inverse depth when
    ``inverse_depth`` (it multiplies the translation column), depth otherwise."""
    input_image = _f(input_image)
    left_cam, right_cam = _f(left_cam), _f(right_cam)
    B, H, W, C = input_image.shape
    R_l, R_r = left_cam[:, 0, :3, :3], right_cam[:, 0, :3, :3]
    t_l, t_r = left_cam[:, 0, :3, 3:4], right_cam[:, 0, :3, 3:4]
    K_l, K_r = left_cam[:, 1, :3, :3], right_cam[:, 1, :3, :3]
    R_l_T = np.transpose(R_l, (0, 2, 1))

    def mv3(m, v):
        return ((m[:, :, 0:1] * v[:, 0:1, :] + m[:, :, 1:2] * v[:, 1:2, :]) + m[:, :, 2:3] * v[:, 2:3, :])

    c_l = -mv3(R_l_T, t_l)
    mat = mm3(K_r, mm3(R_r, mm3(R_l_T, inv3(K_l))))             # (B,3,3)
    vec = mv3(K_r, mv3(R_r, c_l)) + mv3(K_r, t_r)               # (B,3,1)
    grid = get_pixel_grids(H, W).reshape(3, -1)
    px, py = grid[0][None], grid[1][None]
    dep = _f(depth_image).reshape(B, 1, H * W)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        vec = vec * dep if inverse_depth else vec / dep           # (B,3,HW)
        m = mat.reshape(B, 9, 1)
        q0 = ((m[:, 0] * px + m[:, 1] * py) + m[:, 2]) + vec[:, 0]
        q1 = ((m[:, 3] * px + m[:, 4] * py) + m[:, 5]) + vec[:, 1]
        q2 = ((m[:, 6] * px + m[:, 7] * py) + m[:, 8]) + vec[:, 2]
        xw = (q0 / q2).reshape(-1)
        yw = (q1 / q2).reshape(-1)
    if output_mask:
        out, mask = interpolate(input_image, xw, yw, output_mask=True, method=method)
        return out.reshape(B, H, W, C), mask.reshape(B, H, W, 1)
    return interpolate(input_image, xw, yw, method=method).reshape(B, H, W, C)
