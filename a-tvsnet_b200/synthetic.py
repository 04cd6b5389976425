"""Seeded synthetic inputs of SURVEY.md section 8(d): orbit camera rig, smooth N(0,1)
features.  NumPy only (host side); used by bench.py, smoke() and the tests."""
import numpy as np


def orbit_cams(n_views, h, w, depth_num, start=0.05, span=0.4224, angle_deg=4.0):
    """(N,2,4,4) float32 cams at FEATURE resolution (h,w): K = [[.55w,0,w/2],[0,.55w,h/2],[0,0,1]],
    ref R=I,t=0; source i orbits the pivot P=(0,0,5) by +-angle*ceil(i/2) about y (i mod 4 in
    {1,2}) or x; cam[1,3,:] = (start, interval, D, start + D*interval), inverse-depth sweep."""
    cams = np.zeros((n_views, 2, 4, 4), np.float64)
    K = np.array([[0.55 * w, 0, w / 2.0], [0, 0.55 * w, h / 2.0], [0, 0, 1.0]])
    P = np.array([0.0, 0.0, 5.0])
    interval = span / depth_num
    for i in range(n_views):
        th = np.deg2rad(angle_deg * np.ceil(i / 2.0)) * (1.0 if i % 2 == 1 else -1.0)
        c, s = np.cos(th), np.sin(th)
        if i == 0:
            R = np.eye(3)
        elif i % 4 in (1, 2):
            R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        else:
            R = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
        C = P - R.T @ P
        t = -R @ C
        cams[i, 0, :3, :3] = R
        cams[i, 0, :3, 3] = t
        cams[i, 0, 3, 3] = 1.0
        cams[i, 1, :3, :3] = K
        cams[i, 1, 3, :] = (start, interval, depth_num, start + depth_num * interval)
    return cams.astype(np.float32)


def smooth_features(n_views, h, w, f=32, seed=0, sigma=1.5):
    """(N,h,w,F) float32: N(0,1) noise blurred with a separable Gaussian then re-standardised."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_views, h, w, f)).astype(np.float32)
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    for axis in (1, 2):
        pad = [(0, 0)] * 4
        pad[axis] = (r, r)
        xp = np.pad(x, pad, mode='reflect')
        acc = np.zeros_like(x)
        for j, kv in enumerate(k):
            sl = [slice(None)] * 4
            sl[axis] = slice(j, j + x.shape[axis])
            acc += kv * xp[tuple(sl)]
        x = acc
    x -= x.mean()
    x /= x.std()
    return x.astype(np.float32)
