"""Random but reproducible test cases for copy / contraction, shared by the CPU and GPU tests.
Modelled on the reference's own test strategy: tests/contract.cpp enumerates label groups, orders,
conjugations, partitions {OnMaster, OnEveryone, Replicated} and random `from` offsets
(contract.cpp:67-80,:121-142,:323-433); ns_copy_test (dist.h:1919-2116) checks copies on
index-valued tensors."""
import numpy as np

from oracle import oracle as O

LETTERS = "abcdefghijklmnopqrstuvwxyz"
DTYPES = [np.float32, np.float64, np.complex64, np.complex128, np.int32]


def fill(n, dtype, seed):
    """Deterministic, index-valued but non-trivial data (every element distinct)."""
    rng = np.random.default_rng(seed)
    dtype = np.dtype(dtype)
    if dtype.kind == "i":
        return rng.integers(-1000, 1000, size=n).astype(dtype)
    if dtype.kind == "c":
        r = rng.uniform(-1, 1, size=n) + 1j * rng.uniform(-1, 1, size=n)
        return r.astype(dtype)
    return rng.uniform(-1, 1, size=n).astype(dtype)


def random_partition(rng, order, dim, nparts):
    """One of the partition flavours the reference tests use, as an int array [nparts][2][N]."""
    n = len(dim)
    kind = rng.choice(["block", "block", "master", "replicated", "halo", "random"])
    p = np.zeros((nparts, 2, n), dtype=np.int32)
    if nparts == 1 and kind in ("block", "master", "replicated"):
        p[0, 1] = dim
        return p
    if kind == "master":
        p[0, 1] = dim
    elif kind == "replicated":
        p[:, 1] = dim
    elif kind in ("block", "halo"):
        # split nparts over one or two labels
        procs = [1] * n
        rest = nparts
        for k in rng.permutation(n):
            if rest == 1:
                break
            f = int(rng.choice([d for d in range(1, rest + 1) if rest % d == 0]))
            procs[k] = f
            rest //= f
        if rest > 1:
            procs[int(rng.integers(n))] *= rest
        if kind == "block":
            labels = "".join(order[k] for k in rng.permutation(n))
            p = O.basic_partitioning(order, dim, procs, labels, nparts, 1)
        else:
            ext = [int(rng.integers(0, 2)) for _ in range(n)]
            p = O.basic_partitioning_ext(dim, procs, nparts, False, ext)
    else:
        for i in range(nparts):
            size = [int(rng.integers(0, d + 1)) for d in dim]
            if rng.random() < 0.7:
                size = [max(s, 1) for s in size]
            frm = [int(rng.integers(0, d)) for d in dim]
            if int(np.prod(size)) == 0:
                frm, size = [0] * n, [0] * n
            p[i, 0], p[i, 1] = frm, size
    return np.asarray(p, dtype=np.int32)


def random_copy_case(rng, nparts0=None, nparts1=None, max_dim=6, max_nd=5, dtypes=None):
    n0 = int(rng.integers(1, max_nd + 1))
    labs = list(rng.permutation(list(LETTERS)))
    o0 = "".join(labs[:n0])
    dim0 = [int(rng.integers(1, max_dim + 1)) for _ in range(n0)]
    size0 = [int(rng.integers(1, d + 1)) for d in dim0]
    if rng.random() < 0.4:
        size0 = list(dim0)
    from0 = [int(rng.integers(0, d)) for d in dim0]
    # destination labels: all source labels with size>1, maybe the size-1 ones, maybe new ones
    keep = [l for l, s in zip(o0, size0) if s > 1 or rng.random() < 0.5]
    extra = labs[n0:n0 + int(rng.integers(0, 3))]
    o1l = list(keep) + list(extra)
    if not o1l:
        o1l = [labs[n0]]
    o1 = "".join(rng.permutation(o1l))
    dim1, from1 = [], []
    same_dims = rng.random() < 0.5
    for l in o1:
        s = size0[o0.index(l)] if l in o0 else 1
        d = s + int(rng.integers(0, 3))
        if same_dims and l in o0:
            d = dim0[o0.index(l)]
        dim1.append(d)
        from1.append(int(rng.integers(0, d)))
    if nparts0 is None:
        nparts0 = int(rng.integers(1, 5))
    if nparts1 is None:
        nparts1 = int(rng.integers(1, 5))
    p0 = random_partition(rng, o0, dim0, nparts0)
    p1 = random_partition(rng, o1, dim1, nparts1)
    co = int(rng.integers(0, 2))
    copyadd = int(rng.integers(0, 2))
    dts = dtypes or DTYPES
    T = np.dtype(dts[int(rng.integers(len(dts)))])
    conv = {np.dtype(np.float32): np.float64, np.dtype(np.float64): np.float32,
            np.dtype(np.complex64): np.complex128, np.dtype(np.complex128): np.complex64}
    Q = np.dtype(conv[T]) if (T in conv and rng.random() < 0.25) else T
    if T.kind == "c":
        alpha = [1, 1, 0, -1, 2.5 - 0.5j, 0.3 + 1.7j][int(rng.integers(6))]
    elif T.kind == "i":
        alpha = [1, 1, 0, -1, 3][int(rng.integers(5))]
    else:
        alpha = [1, 1, 0, -1, 2.5, 0.3][int(rng.integers(6))]
    return dict(alpha=alpha, p0=p0, o0=o0, from0=from0, size0=size0, dim0=dim0, p1=p1, o1=o1,
                from1=from1, dim1=dim1, co=co, copyadd=copyadd, T=T, Q=Q)


def safe_for_reference(case):
    """The reference translates boxes between the two lattices assuming that a box that wraps around
    in one of them wraps identically in the other (translate_range, dist.h:623-640, after
    intersection_aux keeps a wrapped box whole when one interval spans the dimension,
    dist.h:385-394).  That only holds when the label has the same extent in both tensors; otherwise
    its result contradicts its own checker (dist.h:2044-2047).  Cases outside that are compared
    with the oracle only ("reference corner")."""
    for k, l in enumerate(case["o0"]):
        if l not in case["o1"]:
            continue
        m = case["o1"].index(l)
        if case["dim0"][k] == case["dim1"][m]:
            continue
        if case["from0"][k] + case["size0"][k] > case["dim0"][k]:
            return False
        if case["from1"][m] + case["size0"][k] > case["dim1"][m]:
            return False
        for p, idx, dim in ((case["p0"], k, case["dim0"][k]), (case["p1"], m, case["dim1"][m])):
            for i in range(p.shape[0]):
                if p[i, 0, idx] + p[i, 1, idx] > dim:
                    return False
    return True


def make_copy_data(case, seed, consistent=False):
    """consistent=True: overlapping source parts hold the same global tensor (what a replicated or
    halo partition means).  With inconsistent replicas `Copy` is ill defined: the reference lets the
    last holder win (every holder is copied in turn); this implementation reads each destination
    element from one holder only."""
    v0 = [fill(int(np.prod(case["p0"][i, 1])), case["T"], seed * 1000 + i)
          for i in range(case["p0"].shape[0])]
    if consistent:
        glob = fill(int(np.prod(case["dim0"])), case["T"], seed * 1000 + 999)
        gs = np.asarray(O.get_strides(case["dim0"], case["co"]), dtype=np.int64)
        for i in range(case["p0"].shape[0]):
            if v0[i].size:
                c = (O._local_coords(case["p0"][i, 1], case["co"]) + case["p0"][i, 0]) % \
                    np.asarray(case["dim0"])
                v0[i][:] = glob[(c * gs).sum(axis=1)]
    v1 = [fill(int(np.prod(case["p1"][j, 1])), case["Q"], seed * 1000 + 500 + j)
          for j in range(case["p1"].shape[0])]
    return v0, v1


def make_masks(case, seed, density=0.5):
    """Compatible masks (MaskType = float32) for a copy case: one global mask on the source lattice;
    mask0[i] is its restriction to source part i, mask1[j] holds, for every destination element
    inside the copied range, the mask value of the source element that lands there, and arbitrary
    values elsewhere.  This is the only kind of mask pair the reference accepts (equal counts per
    box, tensor.h:1022-1027).  Nonzero values other than 1 are used on purpose."""
    rng = np.random.default_rng(seed * 31 + 7)
    dim0, dim1 = np.asarray(case["dim0"], dtype=np.int64), np.asarray(case["dim1"], dtype=np.int64)
    o0, o1 = case["o0"], case["o1"]
    n0 = len(o0)
    vol0 = int(np.prod(dim0))
    g = (rng.random(vol0) < density) * rng.choice([1.0, 2.0, -1.0, 0.5], size=vol0)
    g = g.astype(np.float32)
    gs = np.asarray(O.get_strides(case["dim0"], case["co"]), dtype=np.int64)
    mask0 = []
    for i in range(case["p0"].shape[0]):
        c = (O._local_coords(case["p0"][i, 1], case["co"]) + case["p0"][i, 0]) % dim0
        mask0.append(np.ascontiguousarray(g[(c * gs).sum(axis=1)]) if c.shape[0] else
                     np.zeros(0, dtype=np.float32))
    size1 = [int(case["size0"][o0.index(l)]) if l in o0 else 1 for l in o1]
    mask1 = []
    for j in range(case["p1"].shape[0]):
        sj = case["p1"][j, 1]
        n = int(np.prod(sj))
        m = (rng.random(n) < density).astype(np.float32)
        if n:
            c1 = (O._local_coords(sj, case["co"]) + case["p1"][j, 0]) % dim1
            inr = O._in_interval(case["from1"], size1, case["dim1"], c1)
            u1 = (c1 - np.asarray(case["from1"], dtype=np.int64)) % dim1
            c0 = np.zeros((n, n0), dtype=np.int64)
            for k, l in enumerate(o0):
                if l in o1:
                    c0[:, k] = u1[:, o1.index(l)]
            c0 = (c0 + np.asarray(case["from0"], dtype=np.int64)) % dim0
            m[inr] = g[(c0[inr] * gs).sum(axis=1)]
        mask1.append(m)
    return mask0, mask1


def oracle_copy(case, v0, v1, mask0=None, mask1=None):
    out = [x.copy() for x in v1]
    O.copy(case["alpha"], case["p0"], case["o0"], case["from0"], case["size0"], case["dim0"], v0,
           case["p1"], case["o1"], case["from1"], case["dim1"], out, case["co"], case["copyadd"],
           mask0=mask0, mask1=mask1)
    return out


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


# --- contraction ----------------------------------------------------------------------------------

def random_contraction_case(rng, nparts=None, max_dim=4, dtypes=None, full_support=True):
    """Labels split in T/A/B/C groups with 0..2 labels each, shuffled orders, random from offsets
    (contract.cpp:76-80), partitions from the flavours with full support."""
    labs = list(rng.permutation(list(LETTERS)))
    groups = {}
    pos = 0
    for g in "TABC":
        k = int(rng.integers(0, 3))
        groups[g] = labs[pos:pos + k]
        pos += k
    if not (groups["T"] or groups["B"] or groups["C"]):
        groups["B"] = [labs[pos]]
        pos += 1
    size = {l: int(rng.integers(1, max_dim + 1)) for g in groups.values() for l in g}
    o0 = "".join(rng.permutation(groups["T"] + groups["A"] + groups["B"])) if (groups["T"] + groups["A"] + groups["B"]) else ""
    o1 = "".join(rng.permutation(groups["T"] + groups["A"] + groups["C"])) if (groups["T"] + groups["A"] + groups["C"]) else ""
    o_r = "".join(rng.permutation(groups["T"] + groups["B"] + groups["C"]))
    if not o0 or not o1:
        return random_contraction_case(rng, nparts, max_dim, dtypes, full_support)

    def tensor(o):
        sz = [size[l] for l in o]
        frm = [int(rng.integers(0, 2)) for _ in o]
        dim = [s + f for s, f in zip(sz, frm)]
        return sz, frm, dim
    size0, from0, dim0 = tensor(o0)
    size1, from1, dim1 = tensor(o1)
    sizer, fromr, dimr = tensor(o_r)
    P = nparts or int(rng.integers(1, 4))

    kinds = []

    def part(o, dim):
        kind = rng.choice(["master", "replicated", "block"])
        kinds.append(kind)
        n = len(dim)
        p = np.zeros((P, 2, n), dtype=np.int32)
        if kind == "master" or P == 1:
            p[0, 1] = dim
        elif kind == "replicated":
            p[:, 1] = dim
        else:
            k = int(rng.integers(n))
            procs = [1] * n
            procs[k] = P
            p = O.basic_partitioning(o, dim, procs, o[k], P, 1)
        return np.asarray(p, dtype=np.int32)
    dts = dtypes or [np.float32, np.float64, np.complex64, np.complex128]
    T = np.dtype(dts[int(rng.integers(len(dts)))])
    sc = [0, 1, -1, 0.5] if T.kind != "c" else [0, 1, -1, 0.5 - 1.5j]
    p0, p1, pr = part(o0, dim0), part(o1, dim1), part(o_r, dimr)
    beta = sc[int(rng.integers(len(sc)))]
    if kinds[2] == "replicated" and P > 1:
        # with overlapping output parts the reference's in-place beta scaling is not well defined
        beta = [0, 1][int(rng.integers(2))]
    return dict(alpha=sc[int(rng.integers(1, len(sc)))], beta=beta,
                p0=p0, from0=from0, size0=size0, dim0=dim0, o0=o0,
                conj0=bool(rng.integers(2)), p1=p1, from1=from1, size1=size1, dim1=dim1,
                o1=o1, conj1=bool(rng.integers(2)), pr=pr, fromr=fromr, sizer=sizer,
                dimr=dimr, o_r=o_r, co=int(rng.integers(0, 2)), T=T)


def make_contraction_data(case, seed):
    mk = lambda p, s: [fill(int(np.prod(p[i, 1])), case["T"], seed * 1000 + s + i)
                       for i in range(p.shape[0])]
    v0, v1, vr = mk(case["p0"], 0), mk(case["p1"], 100), mk(case["pr"], 200)
    # replicated parts of a tensor must agree (same global tensor everywhere)
    for p, v, o, dim in ((case["p0"], v0, case["o0"], case["dim0"]),
                         (case["p1"], v1, case["o1"], case["dim1"]),
                         (case["pr"], vr, case["o_r"], case["dimr"])):
        glob = fill(int(np.prod(dim)), case["T"], seed * 7 + len(o) + 13 * len(v))
        n = len(dim)
        gs = O.get_strides(dim, case["co"])
        for i in range(p.shape[0]):
            if v[i].size == 0:
                continue
            c = (O._local_coords(p[i, 1], case["co"]) + p[i, 0]) % np.asarray(dim)
            v[i][:] = glob[(c * np.asarray(gs)).sum(axis=1)]
    return v0, v1, vr


def oracle_contraction(case, v0, v1, vr):
    out = [x.copy() for x in vr]
    O.contraction(case["alpha"], case["p0"], case["from0"], case["size0"], case["dim0"], case["o0"],
                  case["conj0"], v0, case["p1"], case["from1"], case["size1"], case["dim1"],
                  case["o1"], case["conj1"], v1, case["beta"], case["pr"], case["fromr"],
                  case["sizer"], case["dimr"], case["o_r"], out, case["co"])
    return out
