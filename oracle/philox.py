"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11) in numpy, and the
dropout-mask convention shared by the oracle and the CUDA epilogues.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws dropout masks from TF-1.12's stateful Philox stream
(``deepctr.layers.core.DNN`` -> ``tf.keras.layers.Dropout(rate, seed=1024+i)``,
called from ``model_zoo/DeepCTR/deepctr.py:129``); that stream cannot be
reproduced outside TF, so the mask is *defined* here (SURVEY.md section 7.3
item 6) and implemented identically in ``mamdr_b200/csrc/philox.cuh``:

    key     = (dropout_seed + layer, global_step & 0xffffffff)
    counter = (e >> 2, 0, 0, 0)     with e = row * n_cols + col
    r       = philox4x32_10(counter, key)[e & 3]
    keep    = r < keep_threshold    keep_threshold = min(2^32-1, floor(keep_prob * 2^32))
    mask    = keep ? 1/keep_prob : 0          (fp32)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  Inputs broadcast; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK32
        hi1, lo1 = p1 >> _S32, p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def keep_threshold(keep_prob):
    return min(0xFFFFFFFF, int(np.floor(float(keep_prob) * 4294967296.0)))


def dropout_random_u32(rows, cols, seed, step):
    """uint32 [rows, cols] of the per-element random words defined in the module docstring."""
    assert cols % 4 == 0, "dropout layers need a width divisible by 4"
    n4 = rows * cols // 4
    r = philox4x32_10(np.arange(n4, dtype=np.uint64), 0, 0, 0, seed, int(step) & 0xFFFFFFFF)
    return np.stack(r, axis=1).reshape(rows, cols)


def dropout_mask(rows, cols, seed, step, rate, dtype=np.float32, use_c=True):
    """Inverted-dropout mask M in {0, 1/keep}^{rows x cols}."""
    keep = 1.0 - float(rate)
    if use_c:
        from . import build as _b
        lib = _b.load()
        if lib is not None:
            assert cols % 4 == 0
            out = np.empty((rows, cols), dtype=np.float32)
            lib.oracle_dropout_mask_f32(rows * cols, int(seed) & 0xFFFFFFFF, int(step) & 0xFFFFFFFF,
                                        keep_threshold(keep), float(np.float32(1.0) / np.float32(keep)),
                                        out.ctypes.data)
            return out.astype(dtype, copy=False)
    r = dropout_random_u32(rows, cols, seed, step)
    scale = (np.float32(1.0) / np.float32(keep)).astype(dtype)
    return np.where(r.astype(np.uint64) < np.uint64(keep_threshold(keep)), scale, dtype(0)).astype(dtype)
