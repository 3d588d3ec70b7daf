"""Loader of the native library. There is no Python or CPU fallback: if the CUDA extension is not
built, importing an operation fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsuperbblas_b200.so")
_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "superbblas_b200: %s is missing; build it with `python -m superbblas_b200.build` "
                "(there is no fallback implementation)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.sbb_last_error.restype = ctypes.c_char_p
        _lib.sbb_version.restype = ctypes.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().sbb_last_error().decode())
