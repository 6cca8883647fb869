"""Helpers shared by the -m gpu parity tests (all calls go through the C-ABI)."""
import ctypes as C

import numpy as np
import torch

from mamdr_b200 import _lib

_CTX = {}


def ctx():
    if "c" not in _CTX:
        _CTX["c"] = _lib.Context(0)
    return _CTX["c"]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
