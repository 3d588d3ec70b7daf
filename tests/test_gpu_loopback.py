"""The cross-rank path (pack -> receiver's arena -> unpack, rounds, arena halves, Add ordering) on ONE
GPU: the ranks are "loopback" communicators of this process (sbb_comm_create_local) driven in phases
through the reference's Request interface -- every rank begins the copy, then every rank completes
it.  Same random cases as the multi-process check (tests/dist_check.py), compared bit for bit with
the oracle.  On a 1-GPU box this is the GPU evidence for SURVEY §8 rows a10 / a14; the flag
signalling between processes is covered by tests/test_dist.py::test_nccl_world (>= 2 GPUs) and by
bench.py's result checks at N > 1."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import superbblas_b200 as sb
from tests import cases as C


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def run_case(torch, comms, case, v0, v1, nc0, nc1, chunk=None):
    world = len(comms)
    gpu = sb.createGpuContext(0)
    d0 = [[torch.from_numpy(x.copy()).cuda() for x in v0[r * nc0:(r + 1) * nc0]] for r in range(world)]
    d1 = [[torch.from_numpy(x.copy()).cuda() for x in v1[r * nc1:(r + 1) * nc1]] for r in range(world)]
    reqs = []
    for r in range(world):  # phase 1: every rank begins (packs, signal, local part)
        reqs.append(sb.copy(case["alpha"], case["p0"], nc0, case["o0"], case["from0"], case["size0"],
                            case["dim0"], d0[r], None, gpu, case["p1"], nc1, case["o1"], case["from1"],
                            case["dim1"], d1[r], None, gpu, case["co"], case["copyadd"], comm=comms[r],
                            request=True))
    for q in reqs:          # phase 2: every rank completes (wait, unpack)
        q.wait()
    sb.sync(gpu)
    return [x.cpu().numpy() for r in range(world) for x in d1[r]]


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_loopback_ranks_match_the_oracle(torch_cuda, world, monkeypatch):
    torch = torch_cuda
    comms = sb.comm_create_local(world)
    rng = np.random.default_rng(9000 + world)
    bad = []
    try:
        for it in range(24):
            nc0, nc1 = int(rng.integers(1, 3)), int(rng.integers(1, 3))
            case = C.random_copy_case(rng, nparts0=world * nc0, nparts1=world * nc1)
            v0, v1 = C.make_copy_data(case, 700 + it, consistent=case["copyadd"] == 0)
            want = C.oracle_copy(case, v0, v1)
            got = run_case(torch, comms, case, v0, v1, nc0, nc1)
            for j, (g, w) in enumerate(zip(got, want)):
                if not C.bits_equal(g, w):
                    bad.append((it, j))
    finally:
        for c in comms:
            c.destroy()
    assert not bad, bad


def test_loopback_redistribution_and_shift_in_rounds(torch_cuda):
    """BASELINE configs 3 and 5 at reduced size over 8 loopback ranks, with a small pipeline chunk so
    that the exchange runs in many rounds: t-slabs -> (z,t) blocks, and a +1 shift in z and t."""
    torch = torch_cuda
    import subprocess
    import sys
    # SBB_CHUNK_BYTES is read once per process: run this part in a process of its own
    code = r'''
import numpy as np, torch, sys
import superbblas_b200 as sb
from tests import cases as C
from tests.test_gpu_loopback import run_case
world = 8
comms = sb.comm_create_local(world)
dim = [8, 8, 8, 16, 4, 3, 4]
pa = sb.basic_partitioning("xyztscn", dim, [1, 1, 1, world, 1, 1, 1], "t", world, 1)
pb = sb.basic_partitioning("xyztscn", dim, [1, 1, 2, 4, 1, 1, 1], "zt", world, 1)
for (p0, p1, from1, add) in [(pa, pb, [0] * 7, 0), (pb, pb, [0, 0, 1, 0, 0, 0, 0], 0),
                             (pb, pb, [0, 0, 0, 1, 0, 0, 0], 1), (pb, pa, [1, 0, 0, 0, 0, 0, 0], 0)]:
    case = dict(alpha=1, p0=p0, o0="xyztscn", from0=[0] * 7, size0=dim, dim0=dim, p1=p1, o1="xyztscn",
                from1=from1, dim1=dim, co=1, copyadd=add, T=np.dtype(np.complex64), Q=np.dtype(np.complex64))
    v0, v1 = C.make_copy_data(case, 3)
    want = C.oracle_copy(case, v0, v1)
    for rep in range(3):   # repeated: the arena halves alternate
        got = run_case(torch, comms, case, v0, v1, 1, 1)
        assert all(C.bits_equal(g, w) for g, w in zip(got, want)), (from1, add, rep)
print("LOOPBACK_OK")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root,
                       env=dict(os.environ, SBB_CHUNK_BYTES="4096", PYTHONPATH=root), timeout=300)
    assert "LOOPBACK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_completing_before_every_rank_began_is_an_error(torch_cuda):
    torch = torch_cuda
    comms = sb.comm_create_local(2)
    gpu = sb.createGpuContext(0)
    dim = [4, 6]
    p0 = sb.basic_partitioning("ab", dim, [2, 1], "a", 2, 1)
    p1 = sb.basic_partitioning("ab", dim, [1, 2], "b", 2, 1)
    x = [torch.zeros(12, device="cuda", dtype=torch.float64) for _ in range(2)]
    y = [torch.zeros(12, device="cuda", dtype=torch.float64) for _ in range(2)]
    q = sb.copy(1, p0, 1, "ab", [0, 0], dim, dim, [x[0]], None, gpu, p1, 1, "ab", [0, 0], dim, [y[0]], None,
                gpu, sb.FastToSlow, sb.Copy, comm=comms[0], request=True)
    with pytest.raises(RuntimeError, match="every rank must begin"):
        q.wait()
    for c in comms:
        c.destroy()


def test_deferred_request_with_host_destination(torch_cuda):
    """A Request on a copy without communicator: the host destination is complete after wait()."""
    torch = torch_cuda
    gpu, cpu = sb.createGpuContext(0), sb.createCpuContext()
    dim0, dim1 = [4, 5, 6], [6, 5, 4]
    case = dict(alpha=2, p0=np.array([[[0] * 3, dim0]], dtype=np.int32), o0="abc", from0=[0] * 3, size0=dim0,
                dim0=dim0, p1=np.array([[[0] * 3, dim1]], dtype=np.int32), o1="cba", from1=[0] * 3, dim1=dim1,
                co=1, copyadd=0, T=np.dtype(np.float64), Q=np.dtype(np.float64))
    v0, v1 = C.make_copy_data(case, 1)
    want = C.oracle_copy(case, v0, v1)
    src = torch.from_numpy(v0[0]).cuda()
    dst = v1[0].copy()
    q = sb.copy(2, case["p0"], 1, "abc", [0] * 3, dim0, dim0, [src], None, gpu, case["p1"], 1, "cba", [0] * 3,
                dim1, [dst], None, cpu, sb.FastToSlow, sb.Copy, request=True)
    q.wait()
    q.wait()  # a completed request can be waited on again
    assert C.bits_equal(dst, want[0])


def test_halo_gather_with_requests_bsr_krylov_pattern(torch_cuda):
    """The hot-path part of the reference's bsr_krylov (bsr.h:2189-2246): the input multi-vector,
    partitioned on z,t, is gathered into halo-extended blocks in the operator's own label order with
    a Request (reorder_tensor_request), the local operator runs, and the result is copied back.
    Here: 8 loopback ranks, halo of width 1 in x,y,z,t with periodic wraparound, label permutation
    "xyztsn" -> "nsxyzt"; the 'operator' is a nearest-neighbour sum evaluated with torch on every
    rank's extended block; the result must equal the same stencil on the global field."""
    torch = torch_cuda
    world = 8
    comms = sb.comm_create_local(world)
    gpu = sb.createGpuContext(0)
    try:
        dim = [8, 8, 8, 8, 2, 3]                      # x y z t s n
        procs = [1, 1, 2, 4, 1, 1]
        px = sb.basic_partitioning("xyztsn", dim, procs, "zt", world, 1)
        ext = sb.basic_partitioning(dim, procs, world, False, [1, 1, 1, 1, 0, 0])   # halo-extended boxes
        perm = [dim[k] for k in (5, 4, 0, 1, 2, 3)]   # dims in "nsxyzt" order
        ext_p = np.array([[[b[0][k] for k in (5, 4, 0, 1, 2, 3)], [b[1][k] for k in (5, 4, 0, 1, 2, 3)]]
                          for b in ext], dtype=np.int32)
        g = torch.Generator(device="cuda").manual_seed(3)
        glob = torch.randn(*reversed(dim), generator=g, device="cuda", dtype=torch.float64)  # [n,s,t,z,y,x]
        one = np.array([[[0] * 6, dim]], dtype=np.int32)
        # scatter the global field onto the ranks (single process source, no communicator needed per rank:
        # every loopback rank takes its own block from the replicated global tensor)
        x = []
        for r in range(world):
            f, s = px[r]
            sl = tuple(slice(int(f[k]), int(f[k] + s[k])) for k in reversed(range(6)))
            x.append(glob[sl].contiguous().view(-1))
        # 1. halo gather with requests: "xyztsn" blocks -> extended "nsxyzt" blocks
        xe = [torch.zeros(int(np.prod(ext_p[r][1])), device="cuda", dtype=torch.float64) for r in range(world)]
        reqs = [sb.copy(1, px, 1, "xyztsn", [0] * 6, dim, dim, [x[r]], None, gpu, ext_p, 1, "nsxyzt", [0] * 6,
                        perm, [xe[r]], None, gpu, sb.FastToSlow, sb.Copy, comm=comms[r], request=True)
                for r in range(world)]
        for q in reqs:
            q.wait()
        # 2. the local operator on every extended block: sum of the 8 neighbours in x,y,z,t (interior only)
        ye = []
        for r in range(world):
            e = ext_p[r][1]                                       # extents in n s x y z t order
            v = xe[r].view(*reversed([int(k) for k in e]))        # [t, z, y, x, s, n]
            acc = torch.zeros_like(v)
            for d in range(4):
                acc += torch.roll(v, 1, dims=d) + torch.roll(v, -1, dims=d)
            ye.append(acc[1:-1, 1:-1].contiguous().view(-1))      # interior in z,t (x,y are not split)
        # 3. copy the result (the operator's own, non-overlapping output partition) back into the
        #    caller's partition and order
        px_p = np.array([[[b[0][k] for k in (5, 4, 0, 1, 2, 3)], [b[1][k] for k in (5, 4, 0, 1, 2, 3)]]
                         for b in px], dtype=np.int32)
        y = [torch.zeros_like(x[r]) for r in range(world)]
        reqs = [sb.copy(1, px_p, 1, "nsxyzt", [0] * 6, perm, perm, [ye[r]], None, gpu, px, 1, "xyztsn", [0] * 6,
                        dim, [y[r]], None, gpu, sb.FastToSlow, sb.Copy, comm=comms[r], request=True)
                for r in range(world)]
        for q in reqs:
            q.wait()
        sb.sync(gpu)
        want = torch.zeros_like(glob)
        for d in (2, 3, 4, 5):                                    # t z y x of [n,s,t,z,y,x]
            want += torch.roll(glob, 1, dims=d) + torch.roll(glob, -1, dims=d)
        for r in range(world):
            f, s = px[r]
            sl = tuple(slice(int(f[k]), int(f[k] + s[k])) for k in reversed(range(6)))
            assert torch.equal(y[r], want[sl].contiguous().view(-1)), r
    finally:
        for c in comms:
            c.destroy()
