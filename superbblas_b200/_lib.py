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
        _lib.sbb_source_hash.restype = ctypes.c_char_p
        # a binary built from other sources than the ones next to it is an error, not a warning:
        # its ABI and kernels may differ from what the tests and the header describe
        if os.path.isdir(os.path.join(_HERE, "csrc")) and not os.environ.get("SBB_ALLOW_STALE_LIB"):
            from .build import source_hash
            have, want = _lib.sbb_source_hash().decode(), source_hash()
            if have != want:
                _lib = None
                raise NativeLibraryMissing(
                    "superbblas_b200: %s was built from other sources (embedded hash %s, sources %s); "
                    "rebuild it with `python -m superbblas_b200.build`" % (LIB_PATH, have[:12], want[:12]))
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().sbb_last_error().decode())
