"""The body of the contraction row kernel (superbblas_b200/csrc/contract_row.hpp) executed on the CPU,
one row per call, through tests/cxx/row_kernel_emul.cpp: random label groups, storage orders, conj
flags, types and alpha/beta against numpy.  The CUDA kernel calls the very same function, so this
checks its indexing and arithmetic without a GPU (the launch itself is covered by the GPU tests once
the kernel is enabled)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import superbblas_b200 as sb
from superbblas_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


class ContractDim(ctypes.Structure):  # sbk_contract_dim (include/superbblas_b200.h)
    _fields_ = [("size", ctypes.c_int), ("s0", ctypes.c_int64), ("s1", ctypes.c_int64),
                ("sr", ctypes.c_int64)]


class ContractDesc(ctypes.Structure):  # sbk_contract_desc
    _fields_ = [("nT", ctypes.c_int), ("nM", ctypes.c_int), ("nN", ctypes.c_int), ("nK", ctypes.c_int),
                ("T", ContractDim * 8), ("M", ContractDim * 8), ("N", ContractDim * 8),
                ("K", ContractDim * 8), ("conj0", ctypes.c_int), ("conj1", ctypes.c_int)]


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not shutil.which("g++") or not os.path.exists(os.path.join(CUDA_INC, "vector_types.h")):
        pytest.skip("needs g++ and the CUDA headers")
    so = str(tmp_path_factory.mktemp("rowk") / "librowk.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I" + CUDA_INC,
                        os.path.join(ROOT, "tests", "cxx", "row_kernel_emul.cpp"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(so)


def _problem(rng, dtype, long_k=False):
    """Random contraction with one big and one small free group (or, long_k, two small free groups
    and a long contraction); every tensor stored in its own random label order (so all strides are
    exercised)."""
    groups = {g: [(g + str(i), int(rng.integers(1, 5))) for i in range(int(rng.integers(0, 3)))]
              for g in "TMNK"}
    if not groups["K"]:
        groups["K"] = [("K0", int(rng.integers(1, 5)))]
    if long_k:
        groups["K"].append(("Kx", int(rng.integers(20, 200))))
    else:
        big = "M" if rng.random() < 0.5 else "N"
        groups[big].append((big + "x", int(rng.integers(5, 40))))  # make one free group the big one
    size = {name: s for g in groups.values() for name, s in g}

    def tensor(labels):
        order = list(rng.permutation(labels)) if labels else []
        stride, acc = {}, 1
        for l in order:  # first label fastest
            stride[l] = acc
            acc *= size[l]
        return order, stride, acc
    l0 = [n for n, _ in groups["T"] + groups["M"] + groups["K"]]
    l1 = [n for n, _ in groups["T"] + groups["N"] + groups["K"]]
    lr = [n for n, _ in groups["T"] + groups["M"] + groups["N"]]
    o0, s0, n0 = tensor(l0)
    o1, s1, n1 = tensor(l1)
    o_r, sr, nr = tensor(lr)
    desc = ContractDesc()
    for g, field, cnt in (("T", desc.T, "nT"), ("M", desc.M, "nM"), ("N", desc.N, "nN"), ("K", desc.K, "nK")):
        setattr(desc, cnt, len(groups[g]))
        for i, (name, s) in enumerate(groups[g]):
            field[i].size = s
            field[i].s0, field[i].s1, field[i].sr = s0.get(name, 0), s1.get(name, 0), sr.get(name, 0)
    desc.conj0, desc.conj1 = int(rng.integers(2)), int(rng.integers(2))

    def data(n):
        x = rng.uniform(-1, 1, n)
        if np.dtype(dtype).kind == "c":
            x = x + 1j * rng.uniform(-1, 1, n)
        return x.astype(dtype)
    v0, v1, vr = data(n0), data(n1), data(nr)
    # numpy reference: tensors as arrays with axes in reversed storage order (first label fastest)
    letters = {name: chr(ord("a") + i) for i, name in enumerate(size)}
    view = lambda v, order: v.reshape([size[l] for l in reversed(order)]) if order else v.reshape(())
    sub = lambda order: "".join(letters[l] for l in reversed(order))
    a, b = view(v0.astype(np.complex128), o0), view(v1.astype(np.complex128), o1)
    if desc.conj0:
        a = a.conj()
    if desc.conj1:
        b = b.conj()
    ref = np.einsum("%s,%s->%s" % (sub(o0), sub(o1), sub(o_r)), a, b).reshape(-1)
    return desc, v0, v1, vr, ref


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_row_kernel_body_against_numpy(emul, dtype):
    rng = np.random.default_rng(2100 + np.dtype(dtype).itemsize + (np.dtype(dtype).kind == "c"))
    DT = {np.float32: sb.F32, np.float64: sb.F64, np.complex64: sb.C64, np.complex128: sb.C128}[dtype]
    cplx = np.dtype(dtype).kind == "c"
    tol = 1e-5 if np.dtype(dtype).itemsize // (2 if cplx else 1) == 4 else 1e-12
    ran = 0
    for it in range(300):
        desc, v0, v1, vr, ref = _problem(rng, dtype)
        if not emul.rowk_eligible(ctypes.byref(desc)):
            continue
        alpha = [1, -1, 0.5 - (1.5j if cplx else 0)][it % 3]
        beta = [0, 1, -0.25 + (0.5j if cplx else 0)][(it // 3) % 3]
        want = alpha * ref + beta * vr.astype(np.complex128)
        if beta == 0:
            vr[:] = np.nan  # beta == 0 must not read the old result
        out = vr.copy()
        rc = emul.rowk_emulate(ctypes.byref(desc), DT, api._scalar(alpha), v0.ctypes.data_as(ctypes.c_void_p),
                               v1.ctypes.data_as(ctypes.c_void_p), api._scalar(beta),
                               out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        if not cplx:
            want = want.real
        err = np.linalg.norm(out.astype(np.complex128) - want) / max(np.linalg.norm(want), 1e-30)
        assert err < tol, (it, err)
        ran += 1
    assert ran >= 150


def test_output_enumeration_is_a_bijection_along_the_smallest_stride(emul):
    """Output enumeration of the generic kernel: every output is visited exactly once and consecutive threads walk the group
    with the smallest result stride."""
    rng = np.random.default_rng(2200)
    emul.rowk_output_index.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_longlong, ctypes.c_longlong,
                                       ctypes.c_longlong, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong)]
    for it in range(200):
        vol = [int(rng.integers(1, 6)) for _ in range(3)]  # T, M, N
        key = [int(x) for x in rng.permutation([1, 7, 50])] if it % 4 else [5, 5, 5]
        min_sr = (ctypes.c_longlong * 3)(*key)
        order = (ctypes.c_int * 3)()
        emul.rowk_output_order(min_sr, order)
        assert sorted(order) == [0, 1, 2]
        assert [key[g] for g in order] == sorted(key)
        if it % 4 == 0:
            assert list(order) == [2, 1, 0]  # ties keep the default enumeration
        seen = set()
        tmn = (ctypes.c_longlong * 3)()
        for idx in range(vol[0] * vol[1] * vol[2]):
            emul.rowk_output_index(order, vol[0], vol[1], vol[2], idx, tmn)
            c = tuple(tmn)
            assert all(0 <= c[g] < vol[g] for g in range(3))
            seen.add(c)
            if idx == 1 and vol[order[0]] > 1:
                assert c[order[0]] == 1  # the fastest group advances first
        assert len(seen) == vol[0] * vol[1] * vol[2]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
def test_dot_kernel_passes_against_numpy(emul, dtype):
    """contract_dot.hpp: partial sums per k slice and 4x4 block (pass 1), ordered reduction with
    alpha/beta (pass 2), for several slice counts."""
    rng = np.random.default_rng(2300 + np.dtype(dtype).itemsize + (np.dtype(dtype).kind == "c"))
    DT = {np.float32: sb.F32, np.float64: sb.F64, np.complex64: sb.C64, np.complex128: sb.C128}[dtype]
    cplx = np.dtype(dtype).kind == "c"
    tol = 1e-5 if np.dtype(dtype).itemsize // (2 if cplx else 1) == 4 else 1e-12
    emul.dotk_emulate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    ran = 0
    for it in range(200):
        desc, v0, v1, vr, ref = _problem(rng, dtype, long_k=True)
        if not emul.dotk_eligible(ctypes.byref(desc)):
            continue
        alpha = [1, -1, 0.5 - (1.5j if cplx else 0)][it % 3]
        beta = [0, 1, -0.25 + (0.5j if cplx else 0)][(it // 3) % 3]
        want = alpha * ref + beta * vr.astype(np.complex128)
        if beta == 0:
            vr[:] = np.nan
        out = vr.copy()
        a, b = api._scalar(alpha), api._scalar(beta)
        rc = emul.dotk_emulate(ctypes.addressof(desc), DT, [64, 1000, 100000][it % 3],
                               ctypes.addressof(a), v0.ctypes.data, v1.ctypes.data, ctypes.addressof(b),
                               out.ctypes.data)
        assert rc == 0
        if not cplx:
            want = want.real
        err = np.linalg.norm(out.astype(np.complex128) - want) / max(np.linalg.norm(want), 1e-30)
        assert err < tol, (it, err)
        ran += 1
    assert ran >= 100
