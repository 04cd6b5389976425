"""Depth-map error metrics with the semantics of /root/reference/atvsnet/eval_errors.py:26-92
(``calc_error``), the evaluation example.py:196,282 runs on its prediction and writes to
``result/error.xlsx``.  Host-side NumPy (as in the reference): evaluation is not on the GPU path.

Order of ``errors`` (float32, 10 + len(inlier_threshold)):
  0 mae | 1 rmse | 2 inverse mae | 3 inverse rmse | 4 log mae | 5 log rmse | 6 scale-invariant log |
  7 abs relative | 8 squared relative | 9 mae in units of (gt range / num_depths) | 10.. inlier ratios
"""
import numpy as np

inlier_thres = [1, 3, 5, 10]
err_metrics_namelist = ['mae', 'rmse', 'inverse_mae', 'inverse_rmse', 'log_mae', 'log_rmse', 'scale_invariant_log',
                        'abs_relative', 'squared_relative', 'mae_normalized']
acc_metrics_namelist = ['inlier_ratios_%d' % t for t in inlier_thres]


def calc_error(depth_predict_in, depth_gt_in, num_depths=100, inlier_threshold=inlier_thres):
    """-> (errors float32[10 + T], infos [num_depths, interval, gt_min, gt_max, thresholds]).
    A pixel counts when both depths are in (0, 1e10) (NaN -> 0 -> invalid); the normalising interval is the
    ground-truth range over its own valid pixels divided by ``num_depths`` (eval_errors.py:35-47)."""
    if depth_predict_in.shape != depth_gt_in.shape:
        raise AssertionError("calc_error: shapes differ %s vs %s" % (depth_predict_in.shape, depth_gt_in.shape))
    pred = np.where(np.isnan(depth_predict_in), 0.0, depth_predict_in).astype(depth_predict_in.dtype)
    gt = np.where(np.isnan(depth_gt_in), 0.0, depth_gt_in).astype(depth_gt_in.dtype)

    gt_ok = (gt > 0.0) & (gt < 1e10)
    gt_vals = gt[gt_ok]
    gt_min, gt_max = gt_vals.min(), gt_vals.max()
    interval = float(gt_max - gt_min) / float(num_depths)

    ok = gt_ok & (pred > 0.0) & (pred < 1e10)
    n = float(ok.sum())
    if not n > 0:
        raise AssertionError("calc_error: no valid pixel")
    # invalid pixels are parked at depth 1 (log = 0, 1/x = 1) and masked out of every sum, in the input dtype
    g = np.where(ok, gt, gt.dtype.type(1.0))
    p = np.where(ok, pred, pred.dtype.type(1.0))
    absd = ok * np.abs(g - p)
    absd2 = absd * absd
    inv = ok * np.abs(1.0 / g - 1.0 / p)
    logd = np.log(g) - np.log(p)
    alog = ok * np.abs(logd)
    alog2 = alog * alog

    e = np.zeros(10 + len(inlier_threshold), dtype=np.float32)
    e[0] = np.sum(absd) / n
    e[1] = np.sqrt(np.float32(np.sum(absd2) / n))
    e[2] = np.sum(inv) / n
    e[3] = np.sqrt(np.float32(np.sum(inv * inv) / n))
    e[4] = np.sum(alog) / n
    mean_log2 = np.sum(alog2) / n
    e[5] = np.sqrt(mean_log2)
    slog = np.sum(ok * logd)
    e[6] = np.sqrt(mean_log2 - (slog * slog / (n * n)))
    e[7] = np.sum(absd / g) / n
    e[8] = np.sum(absd2 / (g * g)) / n
    e[9] = np.sum(absd) / interval / n
    scaled = absd[ok] / interval
    for i, th in enumerate(inlier_threshold):
        e[10 + i] = float(np.sum(scaled < th)) / n
    return e, [num_depths, interval, gt_min, gt_max, inlier_threshold]
