"""Pins the dropout RNG to the Random123 Philox4x32-10 known-answer vectors."""
import numpy as np

from oracle import philox


def _h(r):
    return [int(x) for x in r]


def test_random123_kat():
    assert _h(philox.philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _h(philox.philox4x32_10(f, f, f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _h(philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mask_convention():
    m = philox.dropout_mask(64, 128, 1024, 7, 0.5)
    assert m.dtype == np.float32 and set(np.unique(m)) == {0.0, 2.0}
    assert abs(m.mean() - 1.0) < 0.05
    # element e uses word e&3 of counter e>>2
    r = philox.dropout_random_u32(2, 8, 1024, 7)
    w = philox.philox4x32_10(np.arange(4), 0, 0, 0, 1024, 7)
    assert r[0, 5] == w[1][1] and r[1, 2] == w[2][2]
    # a different step or layer seed gives a different mask
    assert (philox.dropout_mask(64, 128, 1024, 8, 0.5) != m).any()
    assert (philox.dropout_mask(64, 128, 1025, 7, 0.5) != m).any()
    assert philox.keep_threshold(0.5) == 0x80000000 and philox.keep_threshold(1.0) == 0xFFFFFFFF
