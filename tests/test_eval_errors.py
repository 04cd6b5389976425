"""calc_error (a-tvsnet_b200/eval_errors.py) against the ONLY reference-produced known answers the repository
ships: example/{0,1,2}/result/pred.npy evaluated against 0_gt.npy by the reference itself and written to
result/error.xlsx (example.py:196-216).  The xlsx values are committed in tests/golden/reference_error_xlsx.json;
the 2.4 MB arrays are read from /root/reference when it is present (this container), else the test is skipped."""
import importlib.util
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def _mod():
    spec = importlib.util.spec_from_file_location('atvs_eval_errors', os.path.join(ROOT, 'a-tvsnet_b200', 'eval_errors.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize('ex', [0, 1, 2])
def test_calc_error_matches_reference_xlsx(ex):
    pred_p = os.path.join(REF, 'example/%d/result/pred.npy' % ex)
    if not os.path.exists(pred_p):
        pytest.skip('reference example outputs not available')
    m = _mod()
    gold = json.load(open(os.path.join(ROOT, 'tests/golden/reference_error_xlsx.json')))['example/%d' % ex]
    pred = np.squeeze(np.load(pred_p))
    gt = np.squeeze(np.load(os.path.join(REF, 'example/%d/0_gt.npy' % ex)))
    err, info = m.calc_error(pred, gt)
    names = m.err_metrics_namelist + m.acc_metrics_namelist
    assert len(err) == len(names) == 14
    for k, v in zip(names, err):
        assert abs(float(v) - gold[k]) <= 2e-6 * max(1.0, abs(gold[k])), (k, float(v), gold[k])


def test_calc_error_properties():
    m = _mod()
    rng = np.random.default_rng(0)
    gt = rng.uniform(2.0, 12.0, (40, 50)).astype(np.float32)
    e, info = m.calc_error(gt.copy(), gt)
    assert np.all(e[:10] == 0) and np.all(e[10:] == 1.0)
    pred = gt + 0.5
    pred[0, 0] = np.nan          # NaN prediction and non-positive / huge ground truth are excluded
    g2 = gt.copy()
    g2[1, 1] = 0.0
    g2[2, 2] = 1e11
    e, info = m.calc_error(pred, g2)
    assert abs(e[0] - 0.5) < 1e-6 and abs(e[1] - 0.5) < 1e-6
    interval = (g2[(g2 > 0) & (g2 < 1e10)].max() - g2[(g2 > 0) & (g2 < 1e10)].min()) / 100.0
    assert abs(info[1] - interval) < 1e-9 and abs(e[9] - 0.5 / interval) < 1e-3
    with pytest.raises(AssertionError):
        m.calc_error(pred[:5], g2)
