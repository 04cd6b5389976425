"""NumPy fp32 restatement of the depth-map fusion of /root/reference/fusibile:

  fusibile kernel            fusibile/fusibile.cu:138-277 (one thread per reference pixel)
  helpers                    fusibile/fusibile.cu:41-133 (get3Dpoint_cu, project_on_camera, getAngle_cu,
                             disparityDepthConversion_cu2), fusibile/config.h:33-35, 160-189 (dot4, matvecmul4P, matvecmul4)
  host scan                  fusibile/fusibile.cu:279-325 (copy_point_cloud_to_host: row-major, keeps X with x, y, z != 0)
  camera set-up              fusibile/cameraGeometryUtils.h:388, 434-444 (M_inv = P[:, :3]^-1, C4, P_col34, f = K[0,0]);
                             atvsnet/depth_fusion.py:69-91 (P = (K E)[0:3]), :93-112 (fake normals 1/sqrt(3) * [depth > 0]),
                             :183-202 (probability filter), :205-224 (normal_thresh = 360 deg, disp_thresh, num_consistent)

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference binary cannot be built here (OpenCV C++ absent, sm_61 only,
--use_fast_math), so parity is pinned on this restatement ("parity unpinned" against the binary).  Texture fetches are
restated from the CUDA programming guide ("Texture Fetching", linear filtering of an un-normalised float4 texture):
xB = x - 0.5, i = floor(xB), alpha = frac(xB) kept in 9-bit fixed point with 8 fractional bits, texels clamped to the
image; the kernel samples at (x + 0.5, y + 0.5), so xB = x.  A reference pixel samples its own texel exactly."""
import numpy as np

F32 = np.float32


def camera_from_krt(K, R, t):
    """-> dict(P (3,4), M_inv (3,3), C (3,), f) in fp32 (depth_fusion.py:69-91, cameraGeometryUtils.h:388-444)."""
    K, R, t = np.asarray(K, np.float64), np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3)
    P = K @ np.concatenate([R, t[:, None]], axis=1)
    return dict(P=P.astype(F32), M_inv=np.linalg.inv(P[:, :3]).astype(F32), C=(-R.T @ t).astype(F32), f=F32(K[0, 0]))


def fake_normals(depth):
    """depth_fusion.py:93-112: unit-ish normal (1,1,1)/1.732050808 where depth > 0, zero elsewhere -> (H,W,3)."""
    n = np.ones(depth.shape + (3,), F32) / F32(1.732050808)
    return (n * (depth > 0)[..., None].astype(F32)).astype(F32)


def tex2d_linear(img, x, y):
    """float4 texture, un-normalised coordinates, linear filter, clamp: sample at texture coordinate (x + 0.5, y + 0.5)."""
    H, W = img.shape[:2]
    x, y = np.asarray(x, F32), np.asarray(y, F32)
    i, j = np.floor(x), np.floor(y)
    a = (np.floor((x - i) * F32(256.0) + F32(0.5)) / F32(256.0)).astype(F32)
    b = (np.floor((y - j) * F32(256.0) + F32(0.5)) / F32(256.0)).astype(F32)
    i0 = np.clip(i.astype(np.int64), 0, W - 1)
    i1 = np.clip(i.astype(np.int64) + 1, 0, W - 1)
    j0 = np.clip(j.astype(np.int64), 0, H - 1)
    j1 = np.clip(j.astype(np.int64) + 1, 0, H - 1)
    a, b = a[..., None], b[..., None]
    one = F32(1.0)
    return ((one - a) * (one - b) * img[j0, i0] + a * (one - b) * img[j0, i1]
            + (one - a) * b * img[j1, i0] + a * b * img[j1, i1]).astype(F32)


def fuse_reference(ref, normals_depths, images, cams, depth_thresh, normal_thresh, num_consistent, save_texture=True):
    """fusibile.cu:138-277 for reference camera ``ref``.  normals_depths (N,H,W,4) = (nx, ny, nz, depth), images
    (N,H,W,4) or None -> (keep (H,W) bool, coord (H,W,3), normal (H,W,4), texture (H,W,4), number_consistent (H,W))."""
    nd = np.asarray(normals_depths, F32)
    N, H, W, _ = nd.shape
    px, py = np.meshgrid(np.arange(W, dtype=F32), np.arange(H, dtype=F32))
    cr = cams[ref]
    normal = nd[ref]
    depth = normal[..., 3]
    # get3Dpoint_cu: X = M_inv (depth * (x, y, 1) - P[:, 3])
    pt = np.stack([depth * px - cr['P'][0, 3], depth * py - cr['P'][1, 3], depth - cr['P'][2, 3]], axis=-1).astype(F32)
    M = cr['M_inv']
    X = np.stack([(M[r, 0] * pt[..., 0] + M[r, 1] * pt[..., 1]) + M[r, 2] * pt[..., 2] for r in range(3)], axis=-1).astype(F32)
    cons_n = normal.copy()
    cons_t = np.asarray(images[ref], F32).copy() if images is not None else np.zeros((H, W, 4), F32)
    count = np.zeros((H, W), np.int32)
    for i in range(N):
        if i == ref:
            continue
        c = cams[i]
        P = c['P']
        tmp = [((P[r, 0] * X[..., 0] + P[r, 1] * X[..., 1]) + P[r, 2] * X[..., 2]) + P[r, 3] for r in range(3)]
        with np.errstate(all='ignore'):
            u = (tmp[0] / tmp[2]).astype(F32)
            v = (tmp[1] / tmp[2]).astype(F32)
            d = tmp[2].astype(F32)
            inb = (u >= 0) & (u < W) & (v >= 0) & (v < H)
            us, vs = np.where(inb, u, 0).astype(F32), np.where(inb, v, 0).astype(F32)
            tnd = tex2d_linear(nd[i], us, vs)
            dC = cr['C'] - c['C']
            baseline = np.sqrt((dC[0] * dC[0] + dC[1] * dC[1]) + dC[2] * dC[2]).astype(F32)
            dd = (cr['f'] * baseline / d).astype(F32)
            td = (cr['f'] * baseline / tnd[..., 3]).astype(F32)
            ok = inb & ((np.abs(dd - td) / dd) < F32(depth_thresh))
            dot = (tnd[..., 0] * normal[..., 0] + tnd[..., 1] * normal[..., 1]) + tnd[..., 2] * normal[..., 2]
            ang = np.arccos(dot.astype(F32)).astype(F32)
            ang = np.where(np.isnan(ang), F32(0), ang)
            ok &= ang < F32(normal_thresh)
        cons_n = np.where(ok[..., None], cons_n + tnd, cons_n).astype(F32)
        if save_texture and images is not None:
            cons_t = np.where(ok[..., None], cons_t + tex2d_linear(np.asarray(images[i], F32), us, vs), cons_t).astype(F32)
        count += ok.astype(np.int32)
    div = (count.astype(F32) + F32(1.0))[..., None]
    cons_n = (cons_n / div).astype(F32)
    cons_t = (cons_t / div).astype(F32)
    keep = count >= int(num_consistent)
    return keep, X, cons_n, cons_t, count


def fuse(normals_depths, images, cams, depth_thresh=0.01, normal_thresh=np.deg2rad(360.0), num_consistent=2, save_texture=True):
    """all reference cameras in order + the host scan (fusibile.cu:279-325, 425-430) -> (coords (M,3), normals (M,3),
    textures (M,4)), camera-major, row-major inside a camera; points with a zero coordinate are dropped."""
    pts, nrm, tex = [], [], []
    for ref in range(len(cams)):
        keep, X, n, t, _ = fuse_reference(ref, normals_depths, images, cams, depth_thresh, normal_thresh, num_consistent,
                                          save_texture)
        keep = keep & (X[..., 0] != 0) & (X[..., 1] != 0) & (X[..., 2] != 0)
        pts.append(X[keep])
        nrm.append(n[keep][:, :3])
        tex.append(t[keep])
    return np.concatenate(pts), np.concatenate(nrm), np.concatenate(tex)
