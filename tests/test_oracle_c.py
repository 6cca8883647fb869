"""The oracle's C restatement (oracle/c) is bit-identical to the numpy restatement."""
import numpy as np

from oracle import auc as oauc
from oracle import build as obuild
from oracle import philox


def test_c_matches_numpy():
    obuild.build()
    assert obuild.load() is not None
    for rows, cols, seed, step in [(64, 128, 1024, 0), (977, 256, 1025, 12345), (1, 64, 1026, 2 ** 32 + 5)]:
        a = philox.dropout_mask(rows, cols, seed, step, 0.5, use_c=True)
        b = philox.dropout_mask(rows, cols, seed, step, 0.5, use_c=False)
        np.testing.assert_array_equal(a, b)
    a = philox.dropout_mask(8, 8, 3, 4, 0.2, use_c=True)
    np.testing.assert_array_equal(a, philox.dropout_mask(8, 8, 3, 4, 0.2, use_c=False))
    rng = np.random.default_rng(0)
    y = (rng.random(3000) < 0.3).astype(np.float32)
    p = rng.random(3000).astype(np.float32)
    p[:100] = oauc.thresholds(500)[rng.integers(0, 500, 100)].clip(0, 1)
    m1, m2 = oauc.AUC(500), oauc.AUC(500)
    m1.update_state(y, p, use_c=True)
    m2.update_state(y, p, use_c=False)
    np.testing.assert_array_equal(m1.acc, m2.acc)
