"""Streaming 500-threshold ROC-AUC, numpy restatement.  TEST INFRASTRUCTURE.

Follows ``/root/reference/utils/auc.py:110-157`` (threshold table, accumulators),
``utils/auc.py:248-281`` (result: ROC + 'interpolation' summation) and
``utils/metrics_utils.py:297-354`` (tiled strict ``>`` compare, four fp32
``assign_add`` accumulators).  Pinned to the reference's doc-string KAT
(``utils/auc.py:44-56``) by ``tests/test_oracle_auc.py``.
"""
import numpy as np

K_EPSILON = 1e-7  # tf.keras.backend.epsilon()


def thresholds(num_thresholds=500):
    """utils/auc.py:118-126 -- python doubles, later a float32 constant (metrics_utils.py:303)."""
    if num_thresholds <= 1:
        raise ValueError('`num_thresholds` must be > 1.')
    t = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
    return np.asarray([0.0 - K_EPSILON] + t + [1.0 + K_EPSILON], dtype=np.float32)


class AUC(object):
    """State layout matches the CUDA accumulators: rows = (tp, fp, fn, tn), fp32 [4, T]."""

    def __init__(self, num_thresholds=500):
        self.num_thresholds = num_thresholds
        self.thresholds = thresholds(num_thresholds)
        self.acc = np.zeros((4, num_thresholds), dtype=np.float32)

    def reset_states(self):  # utils/auc.py:283-284
        self.acc[...] = 0

    def update_state(self, y_true, y_pred, use_c=True):  # metrics_utils.py:245-354
        if use_c:
            from . import build as _b
            lib = _b.load()
            if lib is not None:
                p = np.ascontiguousarray(y_pred, dtype=np.float32).reshape(-1)
                y = np.ascontiguousarray(y_true, dtype=np.float32).reshape(-1)
                lib.oracle_auc_update(p.ctypes.data, y.ctypes.data, p.size, self.thresholds.ctypes.data,
                                      self.num_thresholds, self.acc.ctypes.data)
                return
        y_pred = np.asarray(y_pred, dtype=np.float32).reshape(1, -1)
        label_pos = np.asarray(y_true, dtype=np.float32).reshape(1, -1).astype(bool)
        pred_pos = y_pred > self.thresholds.reshape(-1, 1)  # strict, fp32 compare
        f32 = np.float32
        self.acc[0] += np.sum(label_pos & pred_pos, axis=1).astype(f32)      # tp
        self.acc[1] += np.sum(~label_pos & pred_pos, axis=1).astype(f32)     # fp
        self.acc[2] += np.sum(label_pos & ~pred_pos, axis=1).astype(f32)     # fn
        self.acc[3] += np.sum(~label_pos & ~pred_pos, axis=1).astype(f32)    # tn

    def result(self):  # utils/auc.py:256-281
        return auc_from_counts(self.acc)


def _div_no_nan(a, b):
    out = np.zeros_like(a)
    np.divide(a, b, out=out, where=(b != 0))
    return out


def auc_from_counts(acc):
    acc = np.asarray(acc, dtype=np.float32)
    tp, fp, fn, tn = acc
    recall = _div_no_nan(tp, tp + fn)
    fp_rate = _div_no_nan(fp, fp + tn)
    heights = (recall[:-1] + recall[1:]) / np.float32(2.0)
    return float(np.sum((fp_rate[:-1] - fp_rate[1:]) * heights, dtype=np.float32))
