"""Host-side constants of the streaming AUC metric (device kernels: csrc/auc.cu, csrc/mlp.cu).

Mirrors ``AUC.__init__`` of ``/root/reference/utils/auc.py:110-126``: ``num_thresholds - 2`` linearly
spaced interior thresholds plus the two epsilon-padded end points, stored as a float32 constant
(``utils/metrics_utils.py:303``).
"""
import numpy as np

K_EPSILON = 1e-7


def thresholds(num_thresholds=500):
    if num_thresholds <= 1:
        raise ValueError('`num_thresholds` must be > 1.')
    inner = [(i + 1) * 1.0 / (num_thresholds - 1) for i in range(num_thresholds - 2)]
    return np.asarray([0.0 - K_EPSILON] + inner + [1.0 + K_EPSILON], dtype=np.float32)
