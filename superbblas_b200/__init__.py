"""superbblas_b200 — B200-native implementation of superbblas's tensor hot path (copy / contraction)
behind the reference's own API.  The product is the C-ABI library (include/superbblas_b200.h,
superbblas_b200/lib/libsuperbblas_b200.so) and the C++ header include/superbblas.h; this package is
the Python mirror of that API used by the tests and by bench.py."""
from .api import *  # noqa: F401,F403
from .api import Context, Comm  # noqa: F401
from ._lib import LIB_PATH, NativeLibraryMissing  # noqa: F401
