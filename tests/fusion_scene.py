"""seeded synthetic scene for the depth-map fusion tests: N cameras of the orbit rig looking at the plane z = 5 (world),
exact depth maps by ray / plane intersection, optional outlier block and dropped (zero-depth) pixels."""
import numpy as np


def make_scene(syn, n_views=4, H=48, W=64, seed=0, outliers=True):
    cams = syn.orbit_cams(n_views, H, W, 8)
    K = cams[:, 1, :3, :3].astype(np.float64)
    R = cams[:, 0, :3, :3].astype(np.float64)
    t = cams[:, 0, :3, 3].astype(np.float64)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    p = np.stack([xs, ys, np.ones_like(xs)], axis=-1)
    depths = np.zeros((n_views, H, W), np.float32)
    rng = np.random.default_rng(seed)
    for i in range(n_views):
        C = -R[i].T @ t[i]
        rays = np.einsum('ij,hwj->hwi', R[i].T @ np.linalg.inv(K[i]), p)
        s = (5.0 - C[2]) / rays[..., 2]
        depths[i] = s.astype(np.float32)
    if outliers:
        depths[1, 10:20, 12:30] *= 1.5            # a block of view 1 disagrees with everybody
        drop = rng.random((n_views, H, W)) < 0.05
        depths[drop] = 0.0                         # probability-filtered pixels
    images = rng.uniform(0, 255, (n_views, H, W, 4)).astype(np.float32)
    images[..., 3] = 0
    return K, R, t, depths, images
