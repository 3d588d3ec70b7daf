"""Multi-rank check, run under torch.distributed.run:

  * --backend gloo (CPU, no GPU): every rank plans its part of random distributed copies with the
    native planner, executes the box operations with the numpy interpreter (tests/plan_exec.py) and
    exchanges the packed messages with torch.distributed send/recv; results are compared with the
    oracle evaluated on the whole problem.  This covers the host logic of the N>1 path.
  * --backend nccl (one GPU per rank): the same cases through superbblas_b200.copy /
    .contraction with a communicator (NCCL send/recv inside the library), compared with the oracle.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import superbblas_b200 as sb  # noqa: E402
from tests import cases as C  # noqa: E402
from tests.plan_exec import run_rank  # noqa: E402


def gloo_exchange(wire, rank, Wt):
    def exchange(send):
        reqs, recv = [], {}
        for peer, (s, r) in sorted(wire.items()):
            if r > 0:
                recv[peer] = torch.zeros(r * (2 if Wt.kind == "c" else 1),
                                         dtype=torch.float64 if Wt.itemsize // (2 if Wt.kind == "c" else 1) == 8
                                         else (torch.int32 if Wt.kind == "i" else torch.float32))
                reqs.append(dist.irecv(recv[peer], src=peer))
        for peer, (s, r) in sorted(wire.items()):
            if s > 0:
                buf = np.ascontiguousarray(send[peer])
                t = torch.from_numpy(buf.view(np.float64 if buf.dtype.itemsize // (2 if buf.dtype.kind == "c" else 1) == 8
                                              else (np.int32 if buf.dtype.kind == "i" else np.float32)).copy())
                reqs.append(dist.isend(t, dst=peer))
        for q in reqs:
            q.wait()
        return {p: v.numpy().view(Wt) for p, v in recv.items()}
    return exchange


def stress(n, rank, world, comm, gpu):
    """Back-to-back exchanges without any host synchronisation in between: a handful of prepared
    copies of very different sizes (so that consecutive exchanges use different round counts and
    arena layouts), chosen at random with the same sequence on every rank, with random host delays
    that differ between ranks (so that ranks run ahead of one another by several exchanges).  The
    results are compared on the device after every exchange."""
    import time
    rng = np.random.default_rng(777)       # shared: which case comes next
    mine = np.random.default_rng(1000 + rank)  # private: my delays
    prepared = []
    sizes = [(4, 4, 4, 8, 2), (8, 8, 8, 16, 12), (16, 16, 16, 32, 12), (6, 10, 4, 8, 3)]
    pz = 2 if world % 2 == 0 else 1
    pt = world // pz
    for si, (lx, ly, lz, lt, inner) in enumerate(sizes):
        dim = [lx, ly, lz * pz, lt * pt, inner]
        pa = sb.basic_partitioning("xyztn", dim, [1, 1, 1, world, 1], "t", world, 1)
        pb = sb.basic_partitioning("xyztn", dim, [1, 1, pz, pt, 1], "zt", world, 1)
        for (p0, p1, from1, add) in [(pa, pb, [0] * 5, 0), (pb, pb, [0, 0, 1, 1, 0], 0), (pb, pa, [1, 0, 0, 0, 0], 1)]:
            case = dict(alpha=1, p0=p0, o0="xyztn", from0=[0] * 5, size0=dim, dim0=dim, p1=p1, o1="xyztn",
                        from1=from1, dim1=dim, co=1, copyadd=add, T=np.dtype(np.complex64),
                        Q=np.dtype(np.complex64))
            v0, v1 = C.make_copy_data(case, 40 + si)
            want = C.oracle_copy(case, v0, v1)
            prepared.append(dict(case=case, src=torch.from_numpy(v0[rank].copy()).cuda(),
                                 init=torch.from_numpy(v1[rank].copy()).cuda(),
                                 want=torch.from_numpy(want[rank].copy()).cuda()))
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    stream = torch.cuda.ExternalStream(sb.get_stream(torch.cuda.current_device()))
    with torch.cuda.stream(stream):
        for it in range(n):
            pcase = prepared[int(rng.integers(len(prepared)))]
            case = pcase["case"]
            dst = pcase["init"].clone()
            if mine.random() < 0.3:
                time.sleep(float(mine.random()) * 2e-3)
            sb.copy(1, case["p0"], 1, case["o0"], case["from0"], case["size0"], case["dim0"], [pcase["src"]],
                    None, gpu, case["p1"], 1, case["o1"], case["from1"], case["dim1"], [dst], None, gpu,
                    case["co"], case["copyadd"], comm=comm)
            bad += (dst.view(torch.int64) != pcase["want"].view(torch.int64)).any().to(torch.int64)
    sb.sync(gpu)
    torch.cuda.synchronize()
    nbad = int(bad.item())
    print("rank %d: stress %d exchanges, %d wrong" % (rank, n, nbad), flush=True)
    return nbad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--cases", type=int, default=30)
    ap.add_argument("--stress", type=int, default=0,
                    help="nccl only: this many back-to-back exchanges of mixed sizes with random "
                         "per-rank delays, every result compared bit for bit (flag protocol, arena halves)")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if args.backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(sb.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = sb.comm_create(bytes(uid.cpu().numpy().tobytes()), world, rank, local)
        gpu = sb.createGpuContext(local)
    else:
        dist.init_process_group("gloo")
    rng = np.random.default_rng(4242)  # the same sequence on every rank
    bad = 0
    for it in range(args.cases):
        if rank == 0 and os.environ.get("SBB_DIST_VERBOSE"):
            print("copy case %d" % it, flush=True)
        nc0, nc1 = int(rng.integers(1, 3)), int(rng.integers(1, 3))
        case = C.random_copy_case(rng, nparts0=world * nc0, nparts1=world * nc1)
        masked = it % 3 == 2  # every third case carries (compatible) masks on both tensors
        v0, v1 = C.make_copy_data(case, 100 + it, consistent=masked or case["copyadd"] == 0)
        m0 = m1 = mm0 = mm1 = None
        if masked:
            m0, m1 = C.make_masks(case, 100 + it)
            mm0, mm1 = m0[rank * nc0:(rank + 1) * nc0], m1[rank * nc1:(rank + 1) * nc1]
        want = C.oracle_copy(case, v0, v1, m0, m1)
        mine0 = [x.copy() for x in v0[rank * nc0:(rank + 1) * nc0]]
        mine1 = [x.copy() for x in v1[rank * nc1:(rank + 1) * nc1]]
        if args.backend == "gloo":
            Wt = case["T"] if (case["copyadd"] == 1 and case["T"] != case["Q"]) else case["Q"]
            zero = np.real(case["alpha"]) == 0 and np.imag(case["alpha"]) == 0
            ops, wire = sb.copy_plan(Wt.itemsize, case["p0"], nc0, case["o0"], case["from0"],
                                     case["size0"], case["dim0"], case["p1"], nc1, case["o1"],
                                     case["from1"], case["dim1"], world, rank, case["co"],
                                     case["copyadd"], zero)
            tmp = None
            if masked and not zero:
                # step 1 of a masked copy (capi.cpp, sbb_copy): mask0 travels to the destination
                # layout as an ordinary float copy with the same geometry
                F = np.dtype(np.float32)
                mops, mwire = sb.copy_plan(4, case["p0"], nc0, case["o0"], case["from0"],
                                           case["size0"], case["dim0"], case["p1"], nc1, case["o1"],
                                           case["from1"], case["dim1"], world, rank, case["co"], 0,
                                           False)
                tmp = [np.full(x.size, np.nan, dtype=np.float32) for x in mine1]
                run_rank(mops, mwire, rank, world, nc0, nc1, mm0, tmp, 1, 0, F, F,
                         gloo_exchange(mwire, rank, F))
            run_rank(ops, wire, rank, world, nc0, nc1, mine0, mine1, case["alpha"], case["copyadd"],
                     case["T"], case["Q"], gloo_exchange(wire, rank, Wt), mask_a=tmp,
                     mask_b=mm1 if masked else None)
            got = mine1
        else:
            d0 = [torch.from_numpy(x).cuda() for x in mine0]
            d1 = [torch.from_numpy(x).cuda() for x in mine1]
            dm0 = [torch.from_numpy(x).cuda() for x in mm0] if masked else None
            dm1 = [torch.from_numpy(x).cuda() for x in mm1] if masked else None
            sb.copy(case["alpha"], case["p0"], nc0, case["o0"], case["from0"], case["size0"],
                    case["dim0"], d0, dm0, gpu, case["p1"], nc1, case["o1"], case["from1"],
                    case["dim1"], d1, dm1, gpu, case["co"], case["copyadd"], comm=comm)
            sb.sync(gpu)
            got = [x.cpu().numpy() for x in d1]
        for j, g in enumerate(got):
            if not C.bits_equal(g, want[rank * nc1 + j]):
                bad += 1
                print("rank %d: copy case %d part %d differs" % (rank, it, j), flush=True)
    if args.backend == "nccl" and args.stress > 0:
        bad += stress(args.stress, rank, world, comm, gpu)
    if args.backend == "nccl":
        # distributed contractions: partitions over `world` ranks, reduction of partial sums
        for it in range(args.cases):
            case = C.random_contraction_case(rng, nparts=world, max_dim=5)
            v0, v1, vr = C.make_contraction_data(case, 300 + it)
            want = C.oracle_contraction(case, v0, v1, vr)
            d0, d1, dr = (torch.from_numpy(x[rank].copy()).cuda() for x in (v0, v1, vr))
            sb.contraction(case["alpha"], case["p0"], case["from0"], case["size0"], case["dim0"], 1,
                           case["o0"], case["conj0"], [d0], gpu, case["p1"], case["from1"],
                           case["size1"], case["dim1"], 1, case["o1"], case["conj1"], [d1], gpu,
                           case["beta"], case["pr"], case["fromr"], case["sizer"], case["dimr"], 1,
                           case["o_r"], [dr], gpu, case["co"], comm=comm)
            sb.sync(gpu)
            g, w = dr.cpu().numpy(), want[rank]
            tol = 1e-12 if case["T"] in (np.dtype(np.float64), np.dtype(np.complex128)) else 1e-5
            if w.size and not np.linalg.norm(g - w) <= tol * max(np.linalg.norm(w), np.sqrt(w.size)):
                bad += 1
                print("rank %d: contraction case %d differs" % (rank, it), flush=True)
    t = torch.tensor([bad], dtype=torch.int64, device="cuda" if args.backend == "nccl" else "cpu")
    dist.all_reduce(t)
    if rank == 0:
        print("DIST_CHECK %s world=%d failures=%d" % (args.backend, world, int(t.item())), flush=True)
    dist.destroy_process_group()
    return 1 if int(t.item()) else 0


if __name__ == "__main__":
    sys.exit(main())
