"""Bug hunting for the host planner without a GPU: random copies (1-8 ranks, 1-2 components per rank, up to 7
labels, Copy/Add, conversions, compatible masks) planned by the C++ planner, executed by the numpy
interpreter tests/plan_exec.py with the messages passed through a dict, compared bit for bit with the
oracle.   python scripts/fuzz_planner.py <seed0> <SBB_CHUNK_BYTES> <seconds>
(400 000 cases over chunk sizes 0..100000 passed on 2026-10-18.)"""
import os, sys, time
os.environ["SBB_CHUNK_BYTES"] = sys.argv[2]
sys.path.insert(0, "/root/repo")
import numpy as np
import superbblas_b200 as sb
from tests import cases as C
from tests.plan_exec import run_rank

def run_planned(case, v0, v1, nranks, m0=None, m1=None):
    P0, P1 = case["p0"].shape[0], case["p1"].shape[0]
    nc0, nc1 = P0 // nranks, P1 // nranks
    alpha = case["alpha"]
    zero = np.real(alpha) == 0 and np.imag(alpha) == 0
    def execute(vin, vout, alpha, copyadd, T, Q, zero, ma=None, mb=None):
        wire_t = T if (copyadd == 1 and T != Q) else Q
        plans = [sb.copy_plan(wire_t.itemsize, case["p0"], nc0, case["o0"], case["from0"], case["size0"], case["dim0"],
                              case["p1"], nc1, case["o1"], case["from1"], case["dim1"], nranks, r, case["co"], copyadd, zero)
                 for r in range(nranks)]
        for r in range(nranks):
            for peer, (s, _) in plans[r][1].items():
                assert plans[peer][1].get(r, (0, 0))[1] == s
        mailbox = {}
        for phase in (0, 1):
            for r in range(nranks):
                ops, wire = plans[r]
                def exchange(send, r=r):
                    if phase == 0:
                        for peer, buf in send.items():
                            mailbox[(r, peer)] = buf.copy()
                        raise StopIteration
                    return {peer: mailbox[(peer, r)] for peer, w in wire.items() if w[1] > 0}
                try:
                    run_rank(ops, wire, r, nranks, nc0, nc1, vin[r * nc0:(r + 1) * nc0], vout[r * nc1:(r + 1) * nc1], alpha, copyadd, T, Q, exchange,
                             mask_a=None if ma is None else ma[r * nc1:(r + 1) * nc1], mask_b=None if mb is None else mb[r * nc1:(r + 1) * nc1])
                except StopIteration:
                    pass
    out = [x.copy() for x in v1]
    tmp = None
    if m0 is not None and not zero:
        F = np.dtype(np.float32)
        tmp = [np.full(x.size, np.nan, dtype=np.float32) for x in v1]
        execute(m0, tmp, 1, 0, F, F, False)
    execute(v0, out, alpha, case["copyadd"], case["T"], case["Q"], zero, tmp, m1)
    return out

seed0 = int(sys.argv[1]); tmax = float(sys.argv[3])
t0 = time.time(); n = 0; bad = 0
seed = seed0
while time.time() - t0 < tmax:
    seed += 1
    rng = np.random.default_rng(1234567 + seed)
    nranks = int(rng.integers(1, 9))
    nc0, nc1 = int(rng.integers(1, 3)), int(rng.integers(1, 3))
    case = C.random_copy_case(rng, nparts0=nranks * nc0, nparts1=nranks * nc1, max_dim=int(rng.integers(2, 10)), max_nd=int(rng.integers(1, 8)))
    if np.prod(case["dim0"], dtype=np.int64) > 200000 or np.prod(case["dim1"], dtype=np.int64) > 200000: continue
    masked = rng.random() < 0.4
    v0, v1 = C.make_copy_data(case, seed, consistent=masked or case["copyadd"] == 0)
    m0 = m1 = None
    if masked:
        m0, m1 = C.make_masks(case, seed, density=float(rng.choice([0.1, 0.5, 0.9])))
        r = rng.random()
        if r < 0.15: m0 = None
        elif r < 0.3: m1 = None
    want = C.oracle_copy(case, v0, v1, m0, m1)
    try:
        got = run_planned(case, v0, v1, nranks, m0, m1)
    except Exception as e:
        print("EXC", seed, repr(e)[:300], case, flush=True); bad += 1; continue
    for j, (g, w) in enumerate(zip(got, want)):
        if not C.bits_equal(g, w):
            print("DIFF", seed, j, nranks, masked, case, flush=True); bad += 1; break
    n += 1
print("done", seed0, n, "bad", bad, round(time.time() - t0, 1))
