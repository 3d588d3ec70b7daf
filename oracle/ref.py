"""TEST INFRASTRUCTURE — ctypes loader for oracle/_ref/libsbref.so (the unmodified reference,
CPU build, behind the C ABI of oracle/ref_capi.cpp).  Same calling convention as oracle.oracle so a
test can run both on identical inputs.  Never imported by the product."""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libsbref.so")
_lib = None

DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2,
      np.dtype(np.complex128): 3, np.dtype(np.int32): 4}


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        # The only unprefixed BLAS in the image is the OpenBLAS bundled with opencv; its own
        # dependencies sit beside it and are not on the loader path, so load them first.
        import glob
        d = os.environ.get(
            "SBREF_BLASDIR",
            "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs")
        for pat in ("libquadmath*", "libgfortran*", "libopenblas*"):
            for f in sorted(glob.glob(os.path.join(d, pat))):
                ctypes.CDLL(f, mode=ctypes.RTLD_GLOBAL)
        _lib = ctypes.CDLL(_PATH)
        _lib.sbref_last_error.restype = ctypes.c_char_p
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().sbref_last_error().decode())


def _ia(x):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _ptrs(v):
    arr = (ctypes.c_void_p * max(len(v), 1))()
    for i, x in enumerate(v):
        arr[i] = x.ctypes.data if x.size else None
    return arr


def _co(co):
    return {0: 0, 1: 1, "SlowToFast": 0, "FastToSlow": 1}[co]


def _ab(x):
    a = (ctypes.c_double * 2)(float(np.real(x)), float(np.imag(x)))
    return a


def copy(alpha, p0, o0, from0, size0, dim0, v0, p1, o1, from1, dim1, v1, co, copyadd,
         mask0=None, mask1=None):
    k = []
    a = [_ia(x) for x in (p0, from0, size0, dim0, p1, from1, dim1)]
    k.extend(a)
    m0 = _ptrs(mask0) if mask0 is not None else None
    m1 = _ptrs(mask1) if mask1 is not None else None
    _check(lib().sbref_copy_masked(
        DT[v0[0].dtype], DT[v1[0].dtype], _ab(alpha), len(o0), a[0][1], len(v0), o0.encode(),
        a[1][1], a[2][1], a[3][1], _ptrs(v0), m0, len(o1), a[4][1], len(v1), o1.encode(), a[5][1],
        a[6][1], _ptrs(v1), m1, _co(co), int(copyadd)))


def contraction(alpha, p0, from0, size0, dim0, o0, conj0, v0, p1, from1, size1, dim1, o1, conj1,
                v1, beta, pr, fromr, sizer, dimr, o_r, vr, co):
    a = [_ia(x) for x in (p0, from0, size0, dim0, p1, from1, size1, dim1, pr, fromr, sizer, dimr)]
    _check(lib().sbref_contraction(
        DT[vr[0].dtype], _ab(alpha), len(o0), a[0][1], a[1][1], a[2][1], a[3][1], len(v0),
        o0.encode(), int(bool(conj0)), _ptrs(v0), len(o1), a[4][1], a[5][1], a[6][1], a[7][1],
        len(v1), o1.encode(), int(bool(conj1)), _ptrs(v1), _ab(beta), len(o_r), a[8][1], a[9][1],
        a[10][1], a[11][1], len(vr), o_r.encode(), _ptrs(vr), _co(co)))


def basic_partitioning(order, dim, procs, dist_labels, nprocs=-1, ncomponents=1):
    n = len(dim)
    P = int(np.prod(procs)) if nprocs < 0 else nprocs
    out = np.zeros((P * ncomponents, 2, n), dtype=np.int32)
    d, p = _ia(dim), _ia(procs)
    _check(lib().sbref_basic_partitioning(
        n, order.encode() if order is not None else None, d[1], p[1],
        dist_labels.encode() if dist_labels is not None else None, nprocs, ncomponents,
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))))
    return out


def basic_partitioning_ext(dim, procs, nprocs=-1, replicate=False, ext_power=None):
    n = len(dim)
    P = int(np.prod(procs)) if nprocs < 0 else nprocs
    out = np.zeros((P, 2, n), dtype=np.int32)
    d, p, e = _ia(dim), _ia(procs), _ia(ext_power if ext_power is not None else [0] * n)
    _check(lib().sbref_basic_partitioning_ext(
        n, d[1], p[1], nprocs, int(replicate), e[1],
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))))
    return out


def partitioning_distributed_procs(order, dim, dist_labels, nprocs):
    n = len(dim)
    out = np.zeros(n, dtype=np.int32)
    d = _ia(dim)
    _check(lib().sbref_partitioning_distributed_procs(
        n, order.encode(), d[1], dist_labels.encode(), nprocs,
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))))
    return [int(x) for x in out]


def make_hole(frm, size, hole_from, hole_size, dim):
    n = len(dim)
    out = np.zeros((3 ** n + 1, 2, n), dtype=np.int32)
    nout = ctypes.c_int(0)
    a = [_ia(x) for x in (frm, size, hole_from, hole_size, dim)]
    _check(lib().sbref_make_hole(n, a[0][1], a[1][1], a[2][1], a[3][1], a[4][1],
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                 ctypes.byref(nout)))
    return [(list(map(int, out[i, 0])), list(map(int, out[i, 1]))) for i in range(nout.value)]


def clear_caches():
    _check(lib().sbref_clear_caches())
