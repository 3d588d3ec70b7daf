"""CPU checks of the checkers bench.py uses on the GPU box: the index-valued field (the reference's
SB_DEBUG mock tensor, dist.h:2022-2063) against brute force and against the oracle's copy, and the
workload bookkeeping (flop counts, partitions) of the two contraction configurations."""
import importlib.util
import os

import numpy as np
import torch

import superbblas_b200 as sb
from tests import cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def as_index(f):
    return (f.real.to(torch.float64) + f.imag.to(torch.float64) * (1 << 22)).numpy().astype(np.int64)


def test_index_field_brute_force():
    dim = [4, 3, 6, 5]
    box = (np.array([1, 0, 4, 2]), np.array([3, 3, 4, 2]))
    shift = [0, 2, 1, 0]
    f = bench.index_field(torch, torch.device("cpu"), box, dim, shift, torch.complex64)
    want = []
    for c3 in range(2):
        for c2 in range(4):
            for c1 in range(3):
                for c0 in range(3):
                    c = [(box[0][k] + [c0, c1, c2, c3][k] - shift[k]) % dim[k] for k in range(4)]
                    want.append(c[0] + dim[0] * (c[1] + dim[1] * (c[2] + dim[2] * c[3])))
    assert (as_index(f) == np.array(want)).all()


def test_index_field_predicts_a_distributed_shift():
    """What bench.py asserts at N > 1: copying the index-valued source with displacement `shift`
    gives, on every part, the index field of that part evaluated with the same shift (oracle copy)."""
    dim, world = [4, 4, 4, 8, 3], 4
    part = sb.basic_partitioning("xyztc", dim, [1, 1, 2, 2, 1], "zt", world, 1)
    for shift in ([0, 0, 1, 0, 0], [0, 0, 0, 1, 0], [1, 0, 0, 0, 0]):
        case = dict(alpha=1, p0=part, o0="xyztc", from0=[0] * 5, size0=dim, dim0=dim, p1=part, o1="xyztc",
                    from1=shift, dim1=dim, co=1, copyadd=0, T=np.dtype(np.complex64), Q=np.dtype(np.complex64))
        v0 = [bench.index_field(torch, torch.device("cpu"), part[r], dim, [0] * 5, torch.complex64).numpy()
              for r in range(world)]
        v1 = [np.zeros_like(x) for x in v0]
        got = C.oracle_copy(case, v0, v1)
        for r in range(world):
            want = bench.index_field(torch, torch.device("cpu"), part[r], dim, shift, torch.complex64).numpy()
            assert C.bits_equal(got[r], want)


def test_workloads():
    w = bench.Workload(2, 8)
    assert w.dimv == [3, 32, 32, 64, 256, 64] and w.scaling == "weak"
    assert w.flop_per_gpu == 8.0 * 64 * 64 * 64 * 3 * 32 ** 3
    for n in (1, 2, 4, 8):
        s = bench.Workload(4, n)
        assert s.dimv == [3, 48, 48, 48, 96, 128] and s.scaling == "strong"
        assert abs(s.flop - 4.1747e12) / 4.1747e12 < 1e-4
        assert s.pz * s.pt == n
