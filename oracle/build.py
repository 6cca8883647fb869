"""Builds the oracle's C restatement (oracle/c/*.c -> oracle/_build/liboracle.so) with gcc.
TEST INFRASTRUCTURE.  ``python -m oracle.build``.  The reference itself is pure Python on
TensorFlow 1.12 (no C/C++ sources), so there is no ``oracle/_ref`` to compile."""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(_HERE, "c", "*.c")))
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-shared", "-fPIC", "-o", LIB] + srcs)
    return LIB


_lib = None


def load():
    """ctypes handle or None when the library has not been built."""
    global _lib
    if _lib is None and os.path.exists(LIB):
        lib = ctypes.CDLL(LIB)
        P, I64, U32, F = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32, ctypes.c_float
        lib.oracle_dropout_mask_f32.argtypes = [I64, U32, U32, U32, F, P]
        lib.oracle_philox_words.argtypes = [I64, U32, U32, P]
        lib.oracle_auc_update.argtypes = [P, P, I64, P, ctypes.c_int32, P]
        _lib = lib
    return _lib


if __name__ == "__main__":
    print(build(force=True))
