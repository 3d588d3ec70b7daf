"""Pins the numpy oracle (oracle/oracle.py) against the REAL reference compiled from
/root/reference (oracle/_ref/libsbref.so): partition generators exactly, copy bit for bit,
contraction within rounding."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import cases as C


def test_partition_known_answers(reflib):
    # the reference's own known answers: tests/dist.cpp:103-125
    for dim, labels, n in [([8, 8, 8, 16], "zt", 8), ([4, 4, 4, 8], "xyzt", 6),
                           ([16, 16, 16, 32], "tzyx", 64), ([3, 5, 7, 2], "xyzt", 12),
                           ([6, 6, 6, 6], "xy", 9), ([48, 48, 48, 96], "zt", 8)]:
        assert O.partitioning_distributed_procs("xyzt", dim, labels, n) == \
            reflib.partitioning_distributed_procs("xyzt", dim, labels, n)


def test_partitioning_random(reflib):
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(1, 6))
        order = "".join(rng.permutation(list("xyztsc"))[:n])
        dim = [int(rng.integers(1, 13)) for _ in range(n)]
        labels = "".join(rng.permutation(list(order))[:int(rng.integers(1, n + 1))])
        nprocs = int(rng.integers(1, 25))
        procs = O.partitioning_distributed_procs(order, dim, labels, nprocs)
        assert procs == reflib.partitioning_distributed_procs(order, dim, labels, nprocs)
        ncomp = int(rng.integers(1, 4))
        P = int(np.prod(procs))
        a = O.basic_partitioning(order, dim, procs, labels, P, ncomp)
        b = reflib.basic_partitioning(order, dim, procs, labels, P, ncomp)
        assert np.array_equal(a, b)
        ext = [int(rng.integers(0, 3)) for _ in range(n)]
        a = O.basic_partitioning_ext(dim, procs, P, False, ext)
        b = reflib.basic_partitioning_ext(dim, procs, P, False, ext)
        assert np.array_equal(a, b)


def test_make_hole_sets(reflib):
    rng = np.random.default_rng(2)
    for _ in range(100):
        n = int(rng.integers(1, 4))
        dim = [int(rng.integers(1, 7)) for _ in range(n)]
        frm = [int(rng.integers(0, d)) for d in dim]
        size = [int(rng.integers(1, d + 1)) for d in dim]
        hf = [int(rng.integers(0, d)) for d in dim]
        hs = [int(rng.integers(0, d + 1)) for d in dim]
        a = O.box_elements(O.make_hole(frm, size, hf, hs, dim), dim)
        b = O.box_elements(reflib.make_hole(frm, size, hf, hs, dim), dim)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(6))
def test_copy_random_bit_exact(reflib, seed):
    rng = np.random.default_rng(100 + seed)
    checked = 0
    for it in range(80):
        case = C.random_copy_case(rng)
        if not C.safe_for_reference(case):
            continue
        checked += 1
        v0, v1 = C.make_copy_data(case, seed * 100 + it)
        want = [x.copy() for x in v1]
        reflib.copy(case["alpha"], case["p0"], case["o0"], case["from0"], case["size0"],
                    case["dim0"], v0, case["p1"], case["o1"], case["from1"], case["dim1"], want,
                    case["co"], case["copyadd"])
        got = C.oracle_copy(case, v0, v1)
        for j, (g, w) in enumerate(zip(got, want)):
            assert C.bits_equal(g, w), (seed, it, j, case)
    assert checked >= 30


@pytest.mark.parametrize("seed", range(4))
def test_masked_copy_random_bit_exact(reflib, seed):
    """Masked copies (MaskType masks on both tensors, tensor.h:1022-1027): the reference with
    compatible masks against the oracle, bit for bit."""
    rng = np.random.default_rng(300 + seed)
    checked = 0
    for it in range(80):
        case = C.random_copy_case(rng)
        if not C.safe_for_reference(case):
            continue
        checked += 1
        v0, v1 = C.make_copy_data(case, seed * 100 + it, consistent=True)
        m0, m1 = C.make_masks(case, seed * 100 + it, density=[0.5, 0.1, 0.9][it % 3])
        want = [x.copy() for x in v1]
        reflib.copy(case["alpha"], case["p0"], case["o0"], case["from0"], case["size0"],
                    case["dim0"], v0, case["p1"], case["o1"], case["from1"], case["dim1"], want,
                    case["co"], case["copyadd"], mask0=m0, mask1=m1)
        got = C.oracle_copy(case, v0, v1, m0, m1)
        for j, (g, w) in enumerate(zip(got, want)):
            assert C.bits_equal(g, w), (seed, it, j, case)
    assert checked >= 30


@pytest.mark.parametrize("seed", range(4))
def test_contraction_random(reflib, seed):
    rng = np.random.default_rng(200 + seed)
    for it in range(40):
        case = C.random_contraction_case(rng)
        v0, v1, vr = C.make_contraction_data(case, seed * 100 + it)
        want = [x.copy() for x in vr]
        reflib.contraction(case["alpha"], case["p0"], case["from0"], case["size0"], case["dim0"],
                           case["o0"], case["conj0"], v0, case["p1"], case["from1"], case["size1"],
                           case["dim1"], case["o1"], case["conj1"], v1, case["beta"], case["pr"],
                           case["fromr"], case["sizer"], case["dimr"], case["o_r"], want,
                           case["co"])
        got = C.oracle_contraction(case, v0, v1, vr)
        tol = 1e-12 if case["T"] in (np.dtype(np.float64), np.dtype(np.complex128)) else 1e-5
        for g, w in zip(got, want):
            if w.size == 0:
                continue
            scale = max(np.linalg.norm(w), 1e-30)
            assert np.linalg.norm(g - w) <= tol * max(scale, np.sqrt(w.size)), (seed, it, case)


def test_config1_shapes(reflib):
    """BASELINE config 1: copy "xyztsc"->"cstzyx" and the contraction over s,c on 8^3x16."""
    L, Lt = 8, 16
    dim0, dim1 = [L, L, L, Lt, 4, 3], [3, 4, Lt, L, L, L]
    v0 = C.fill(int(np.prod(dim0)), np.complex128, 11)
    p0, p1 = np.array([[[0] * 6, dim0]]), np.array([[[0] * 6, dim1]])
    for co in ("FastToSlow", "SlowToFast"):
        a, b = [np.zeros_like(v0)], [np.zeros_like(v0)]
        O.copy(1, p0, "xyztsc", [0] * 6, dim0, dim0, [v0], p1, "cstzyx", [0] * 6, dim1, a, co, 0)
        reflib.copy(1, p0, "xyztsc", [0] * 6, dim0, dim0, [v0], p1, "cstzyx", [0] * 6, dim1, b, co,
                    0)
        assert C.bits_equal(a[0], b[0])
    v1 = C.fill(int(np.prod(dim0)), np.complex128, 12)
    dimr = [L, L, L, Lt]
    pr = np.array([[[0] * 4, dimr]])
    a, b = [np.zeros(int(np.prod(dimr)), np.complex128)], [np.zeros(int(np.prod(dimr)), np.complex128)]
    args = (1, p0, [0] * 6, dim0, dim0, "xyztsc", True, [v0], p0, [0] * 6, dim0, dim0, "xyztsc",
            False, [v1], 0, pr, [0] * 4, dimr, dimr, "xyzt")
    O.contraction(*args, a, "FastToSlow")
    reflib.contraction(*args, b, "FastToSlow")
    assert np.linalg.norm(a[0] - b[0]) <= 1e-13 * np.linalg.norm(b[0])
