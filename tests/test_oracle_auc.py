"""Pins the oracle's AUC to the reference's only golden vector for this path:
the doc-string example of /root/reference/utils/auc.py:44-56."""
import numpy as np

from oracle import auc as oauc


def test_docstring_known_answer():
    m = oauc.AUC(num_thresholds=3)
    m.update_state([0, 0, 1, 1], [0, 0.5, 0.3, 0.9])
    np.testing.assert_array_equal(m.thresholds, np.asarray([-1e-7, 0.5, 1 + 1e-7], dtype=np.float32))
    np.testing.assert_array_equal(m.acc[0], [2, 1, 0])   # tp
    np.testing.assert_array_equal(m.acc[1], [2, 0, 0])   # fp
    np.testing.assert_array_equal(m.acc[2], [0, 1, 2])   # fn
    np.testing.assert_array_equal(m.acc[3], [0, 2, 2])   # tn
    assert abs(m.result() - 0.75) < 1e-7


def test_thresholds_500():
    t = oauc.thresholds(500)
    assert t.shape == (500,) and t.dtype == np.float32
    assert t[0] == np.float32(-1e-7) and t[-1] == np.float32(1 + 1e-7)
    assert t[1] == np.float32(1.0 / 499) and np.all(np.diff(t) > 0)


def test_streaming_equals_single_shot_and_sklearn():
    rng = np.random.default_rng(0)
    y = (rng.random(5000) < 0.3).astype(np.float32)
    p = np.clip(0.3 * y + rng.random(5000) * 0.7, 0, 1).astype(np.float32)
    a, b = oauc.AUC(500), oauc.AUC(500)
    a.update_state(y, p)
    for s in range(0, 5000, 1024):
        b.update_state(y[s:s + 1024], p[s:s + 1024])
    np.testing.assert_array_equal(a.acc, b.acc)
    from sklearn.metrics import roc_auc_score
    assert abs(a.result() - roc_auc_score(y, p)) < 2e-3
    a.reset_states()
    assert a.acc.sum() == 0 and a.result() == 0.0


def test_edge_predictions_exactly_zero_and_one():
    m = oauc.AUC(500)
    m.update_state([0, 1, 0, 1], [0.0, 1.0, 1.0, 0.0])
    # p == 0 is above the -eps threshold only; p == 1 is above every threshold but the last
    assert m.acc[0][0] == 2 and m.acc[0][-1] == 0 and m.acc[0][-2] == 1
    assert abs(m.result() - 0.5) < 1e-6
