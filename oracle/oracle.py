"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's tensor hot path.

This module is the *oracle* for the parity tests.  It is not product code: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` leg may import it.  It restates, element
by element, what `superbblas::copy` and `superbblas::contraction` compute; it does not follow the
reference's implementation (index vectors, pack/unpack, GEMM lowering) but its *semantics*:

* copy        : the reference's own analytic checker `ns_copy_test::test_copy_check`
                (/root/reference/include/superbblas/dist.h:2022-2063), the zero-fill rule
                (dist.h:2338-2340, :2356-2382), the alpha==0 rule (dist.h:2350-2353, :2383) and the
                element arithmetic of the CPU loops (copy_n.h:148-300: `w = alpha*v`, `w += alpha*v`).
* contraction : the brute-force contraction of the reference test (tests/contract.cpp:289-321) plus
                `vr = alpha*sum + beta*vr` with beta applied by a `copy` first (dist.h:3145) and the
                product added afterwards (dist.h:3184), replicated inputs counted once
                (dist.h:3001-3028).
* partitions  : basic_partitioning (dist.h:3393-3465, :3477-3509), partitioning_distributed_procs
                (dist.h:3318-3383), factors_2_3 (dist.h:3268-3310), make_hole (dist.h:3750-3825).

Parity pinned: tests/test_oracle_vs_reference.py checks every function here against the real
reference (oracle/_ref/libsbref.so, built from /root/reference by oracle/Makefile) and against the
golden vectors in tests/golden/ (generated from that same library by tests/golden/generate.py).

Conventions: a partition is an int array [P][2][N] (`from`, `size`); tensor data is a list with one
flat numpy array per part (length = prod(size)); `co` is "FastToSlow" (first label fastest) or
"SlowToFast".
"""
import numpy as np

SlowToFast, FastToSlow = 0, 1
Copy, Add = 0, 1

DTYPES = {0: np.float32, 1: np.float64, 2: np.complex64, 3: np.complex128, 4: np.int32}


def _co(co):
    if co in (FastToSlow, "FastToSlow"):
        return FastToSlow
    if co in (SlowToFast, "SlowToFast"):
        return SlowToFast
    raise ValueError("invalid coordinate order")


def get_strides(dim, co):
    """tensor.h:282-299"""
    dim = [int(d) for d in dim]
    n = len(dim)
    s = [1] * n
    if _co(co) == SlowToFast:
        for i in range(n - 2, -1, -1):
            s[i] = s[i + 1] * dim[i + 1]
    else:
        for i in range(1, n):
            s[i] = s[i - 1] * dim[i - 1]
    return s


def _local_coords(size, co):
    """All coordinates of a local box, as an int64 array [vol][N], in memory order."""
    size = [int(s) for s in size]
    vol = int(np.prod(size, dtype=np.int64)) if len(size) else 1
    idx = np.arange(vol, dtype=np.int64)
    strides = get_strides(size, co)
    c = np.empty((vol, len(size)), dtype=np.int64)
    for k, (st, sz) in enumerate(zip(strides, size)):
        c[:, k] = (idx // st) % max(sz, 1)
    return c


def _in_interval(frm, size, dim, c):
    """tensor.h:251-259 (periodic interval membership), vectorised over rows of c."""
    ok = np.ones(c.shape[0], dtype=bool)
    for k in range(c.shape[1]):
        f, s, d = int(frm[k]), int(size[k]), int(dim[k])
        x = c[:, k]
        ok &= ((f <= x) & (x < f + s)) | ((f <= x + d) & (x + d < f + s))
    return ok


def check_copy_args(o0, size0, dim0, o1, dim1):
    """check_isomorphic (tensor.h:495-511): labels unique, size<=dim, every o0 label with size>1
    present in o1 with enough room."""
    if len(set(o0)) != len(o0) or len(set(o1)) != len(o1):
        raise RuntimeError("Invalid copy operation")
    for k, l in enumerate(o0):
        if size0[k] > dim0[k]:
            raise RuntimeError("Invalid copy operation")
        if l in o1:
            if size0[k] > dim1[o1.index(l)]:
                raise RuntimeError("Invalid copy operation")
        elif size0[k] > 1:
            raise RuntimeError("Invalid copy operation")


def _scale(alpha, x, T):
    """alpha*x evaluated in type T with separately rounded real operations (no FMA), as the
    reference's C loops compiled with -ffp-contract=off (copy_n.h:166)."""
    T = np.dtype(T)
    if T.kind == "c":
        R = np.float32 if T == np.complex64 else np.float64
        ar, ai = R(np.real(alpha)), R(np.imag(alpha))
        xr, xi = x.real.astype(R), x.imag.astype(R)
        out = np.empty(x.shape, dtype=T)
        out.real = ar * xr - ai * xi
        out.imag = ar * xi + ai * xr
        return out
    a = T.type(np.real(alpha)) if T.kind != "i" else T.type(int(np.real(alpha)))
    return (a * x).astype(T)


def copy(alpha, p0, o0, from0, size0, dim0, v0, p1, o1, from1, dim1, v1, co, copyadd,
         mask0=None, mask1=None):
    """superbblas::copy (dist.h:3583) on P0 source parts and P1 destination parts held in one
    process.  v1 arrays are modified in place."""
    co = _co(co)
    p0 = np.asarray(p0, dtype=np.int64).reshape(len(v0), 2, len(o0))
    p1 = np.asarray(p1, dtype=np.int64).reshape(len(v1), 2, len(o1))
    check_copy_args(o0, size0, dim0, o1, dim1)
    n0, n1 = len(o0), len(o1)
    # size of the copied range in destination order; labels missing in o0 get 1
    size1 = [int(size0[o0.index(l)]) if l in o0 else 1 for l in o1]
    T = v0[0].dtype if len(v0) else np.dtype(np.float64)
    alpha_is_zero = (np.real(alpha) == 0 and np.imag(alpha) == 0)
    alpha_is_one = (np.real(alpha) == 1 and np.imag(alpha) == 0)
    if copyadd == Add and alpha_is_zero:
        return
    strides0 = [get_strides(p0[i, 1], co) for i in range(len(v0))]
    for j in range(len(v1)):
        fj, sj = p1[j, 0], p1[j, 1]
        if int(np.prod(sj)) == 0:
            continue
        Q = v1[j].dtype
        cl = _local_coords(sj, co)                      # local coords, memory order
        c1 = (cl + fj) % np.asarray(dim1, dtype=np.int64)  # global coords
        inr = _in_interval(from1, size1, dim1, c1)
        if mask1 is not None and mask1[j] is not None:
            inr &= (mask1[j] != 0)
        sel = np.nonzero(inr)[0]
        if sel.size == 0:
            continue
        # range-relative coordinate, then source global coordinate
        u1 = (c1[sel] - np.asarray(from1, dtype=np.int64)) % np.asarray(dim1, dtype=np.int64)
        c0 = np.zeros((sel.size, n0), dtype=np.int64)
        for k, l in enumerate(o0):
            if l in o1:
                c0[:, k] = u1[:, o1.index(l)]
        c0 = (c0 + np.asarray(from0, dtype=np.int64)) % np.asarray(dim0, dtype=np.int64)
        if alpha_is_zero:
            v1[j][sel] = 0
            continue
        covered = np.zeros(sel.size, dtype=bool)
        for i in range(len(v0)):
            fi, si = p0[i, 0], p0[i, 1]
            if int(np.prod(si)) == 0:
                continue
            hit = _in_interval(fi, si, dim0, c0)
            covered |= hit  # coverage is geometric (has_full_support, dist.h:666): masks do not uncover
            l0 = (c0[hit] - fi) % np.asarray(dim0, dtype=np.int64)
            src_idx = (l0 * np.asarray(strides0[i], dtype=np.int64)).sum(axis=1)
            if mask0 is not None and mask0[i] is not None:
                keep = mask0[i][src_idx] != 0
                hidx = np.nonzero(hit)[0][keep]
                hit = np.zeros_like(hit)
                hit[hidx] = True
                src_idx = src_idx[keep]
            if not hit.any():
                continue
            x = v0[i][src_idx]
            if not alpha_is_one:
                x = _scale(alpha, x, T)
            dst_idx = sel[hit]
            if copyadd == Copy:
                v1[j][dst_idx] = x.astype(Q)
            else:
                W = np.result_type(Q, T)
                v1[j][dst_idx] = (v1[j][dst_idx].astype(W) + x.astype(W)).astype(Q)
            covered |= hit
        if copyadd == Copy and not covered.all():
            # source without full support: uncovered destination elements inside the range are zeroed
            v1[j][sel[~covered]] = 0


def _gather_global(p, o, frm, size, dim, v, co):
    """Dense array of the range [frm, frm+size) in label order `o` (C-order axes = o),
    taking each element from the first part that holds it (dist.h:3001-3028)."""
    p = np.asarray(p, dtype=np.int64).reshape(len(v), 2, len(o))
    n = len(o)
    size = [int(s) for s in size]
    vol = int(np.prod(size, dtype=np.int64)) if n else 1
    out = np.zeros(vol, dtype=v[0].dtype)
    u = _local_coords(size, SlowToFast)
    c = (u + np.asarray(frm, dtype=np.int64)) % np.asarray(dim, dtype=np.int64)
    done = np.zeros(vol, dtype=bool)
    for i in range(len(v)):
        fi, si = p[i, 0], p[i, 1]
        if int(np.prod(si)) == 0:
            continue
        hit = _in_interval(fi, si, dim, c) & ~done
        if not hit.any():
            continue
        l = (c[hit] - fi) % np.asarray(dim, dtype=np.int64)
        idx = (l * np.asarray(get_strides(si, co), dtype=np.int64)).sum(axis=1)
        out[hit] = v[i][idx]
        done |= hit
    if not done.all():
        raise RuntimeError("contraction operand without full support on the contracted range")
    return out.reshape(size)


def contraction(alpha, p0, from0, size0, dim0, o0, conj0, v0, p1, from1, size1, dim1, o1, conj1, v1,
                beta, pr, fromr, sizer, dimr, o_r, vr, co):
    """superbblas::contraction (dist.h:3701).  vr arrays are modified in place."""
    co = _co(co)
    # label classes and consistency (tensor.h:1297-1354, dist.h:3122)
    sizes = {}
    for o, s in ((o0, size0), (o1, size1), (o_r, sizer)):
        if len(set(o)) != len(o):
            raise RuntimeError("repeated label")
        for l, x in zip(o, s):
            if sizes.setdefault(l, int(x)) != int(x):
                raise RuntimeError("some dimension does not match")
    for l in o0:
        if l not in o1 and l not in o_r:
            raise RuntimeError("o0 has unmatched dimensions")
    for l in o1:
        if l not in o0 and l not in o_r:
            raise RuntimeError("o1 has unmatched directions")
    for l in o_r:
        if l not in o0 and l not in o1:
            raise RuntimeError("o_r has unmatched dimensions")
    T = vr[0].dtype
    # 1) vr <- beta*vr on the output range (dist.h:3145: copy(beta, vr -> vr)).  Every part scales
    #    its own elements; for partitions without overlaps that is exactly what the reference's
    #    in-place copy does (with overlapping output parts the reference's in-place copy reads
    #    already-scaled replicas, which is not a defined result, so tests use beta in {0,1} there).
    pr_parts = np.asarray(pr, dtype=np.int64).reshape(len(vr), 2, len(o_r))
    for j in range(len(vr)):
        one = pr_parts[j:j + 1]
        src = [vr[j].copy()]
        copy(beta, one, o_r, fromr, sizer, dimr, src, one, o_r, fromr, dimr, [vr[j]], co, Copy)
    # 2) dense contraction of the two ranges in extended precision of the type
    W = np.complex128 if np.dtype(T).kind == "c" else np.float64
    X0 = _gather_global(p0, o0, from0, size0, dim0, v0, co).astype(W)
    X1 = _gather_global(p1, o1, from1, size1, dim1, v1, co).astype(W)
    if conj0:
        X0 = np.conj(X0)
    if conj1:
        X1 = np.conj(X1)
    letters = {}
    pool = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    for l in o0 + o1 + o_r:
        letters.setdefault(l, pool[len(letters)])
    sub = "%s,%s->%s" % ("".join(letters[l] for l in o0), "".join(letters[l] for l in o1),
                         "".join(letters[l] for l in o_r))
    R = np.einsum(sub, X0, X1)
    R = (W(alpha) * R).astype(T).reshape(-1)
    # 3) vr += R on the output range, in every destination part that owns the element (dist.h:3184)
    n = len(o_r)
    pr_ = np.asarray(pr, dtype=np.int64).reshape(len(vr), 2, n)
    pR = np.zeros((1, 2, n), dtype=np.int64)
    pR[0, 1] = [int(s) for s in sizer]
    # R is laid out SlowToFast over sizer; re-express it in `co`
    if co == FastToSlow and n > 0:
        R = R.reshape([int(s) for s in sizer]).transpose(list(range(n - 1, -1, -1))).reshape(-1)
    copy(1, pR, o_r, [0] * n, sizer, sizer, [np.ascontiguousarray(R)], pr_, o_r, fromr, dimr, vr,
         co, Add)


# ---------------------------------------------------------------------------------------------
# Partitioning helpers
# ---------------------------------------------------------------------------------------------

def _factors_2_3(number):
    """dist.h:3268-3310: approximate `number` by 2^a*3^b >= 0.75*number."""
    if number == 0:
        raise RuntimeError("unsupported value")
    two = three = 0
    value = 1
    rem = number
    while rem % 2 == 0:
        two += 1; rem //= 2; value *= 2
    while rem % 3 == 0:
        three += 1; rem //= 3; value *= 3
    while rem >= 3:
        three += 1; rem //= 3; value *= 3
    if rem >= 2:
        two += 1; rem //= 2; value *= 2
    while three > 0 and value * 4 // 3 <= number:
        three -= 1; two += 2; value = value * 4 // 3
    return value


def partitioning_distributed_procs(order, dim, dist_labels, nprocs):
    """dist.h:3318-3383"""
    n = len(order)
    if len(dim) != n:
        raise RuntimeError("order length mismatch")
    p = [1] * n
    dist_perm = [order.index(l) for l in dist_labels if l in order and dim[order.index(l)] > 1]
    vol = int(np.prod([int(d) for d in dim], dtype=np.int64)) if n else 1
    if not dist_perm or vol == 0 or nprocs <= 1:
        return p
    pf = [1] * len(dist_perm)
    vol_p = 1
    nprocs_v = _factors_2_3(nprocs)
    while True:
        # dimensions sorted by local size, largest first (stable selection like the reference)
        perm = list(range(len(dist_perm)))
        for j in range(len(perm)):
            large_i, large_val = j, dim[dist_perm[perm[j]]] // pf[perm[j]]
            for i in range(j + 1, len(perm)):
                val = dim[dist_perm[perm[i]]] // pf[perm[i]]
                if large_val < val:
                    large_i, large_val = i, val
            perm[j], perm[large_i] = perm[large_i], perm[j]
        applied = False
        for j in range(len(perm)):
            for f in (3, 2):
                if nprocs_v % (vol_p * f) == 0:
                    pf[perm[j]] *= f
                    vol_p *= f
                    applied = True
                    break
            if applied:
                break
        if not applied:
            break
    for i, k in enumerate(dist_perm):
        p[k] = pf[i]
    return p


def _index2coor(index, dim, stride):
    return [(index // st) % d for st, d in zip(stride, dim)]


def basic_partitioning(order, dim, procs, dist_labels, nprocs=-1, ncomponents=1):
    """dist.h:3393-3465.  Returns an int array [nprocs*ncomponents][2][N]."""
    n = len(dim)
    dim = [int(d) for d in dim]
    procs = [int(d) for d in procs]
    vol_procs = int(np.prod(procs, dtype=np.int64)) if n else 1
    if order is not None and dist_labels is not None:
        if len(order) != n:
            raise RuntimeError("basic_partitioning: invalid `order`")
        perm = [order.index(l) for l in dist_labels if l in order]
        perm += [i for i in range(n) if order[i] not in dist_labels]
        if len(perm) != n:
            raise RuntimeError("wtf")
    else:
        perm = list(range(n))
    P = vol_procs if nprocs < 0 else nprocs
    fs = np.zeros((P * ncomponents, 2, n), dtype=np.int32)
    procs_perm = [procs[k] for k in perm]
    stride_perm = get_strides(procs_perm, SlowToFast)
    for rank in range(vol_procs):
        cproc = _index2coor(rank, procs_perm, stride_perm)
        frm = [0] * n
        size = [0] * n
        for i in range(n):
            k = perm[i]
            size[k] = dim[k] // procs_perm[i] + (1 if dim[k] % procs_perm[i] > cproc[i] else 0)
            frm[k] = 0 if size[k] == dim[k] else \
                dim[k] // procs_perm[i] * cproc[i] + min(cproc[i], dim[k] % procs_perm[i])
        if int(np.prod(size, dtype=np.int64)) == 0:
            frm = [0] * n
            size = [0] * n
        if ncomponents == 1:
            fs[rank, 0], fs[rank, 1] = frm, size
        else:
            sub = basic_partitioning(
                order, size, partitioning_distributed_procs(order, size, dist_labels, ncomponents),
                dist_labels, ncomponents)
            for c in range(ncomponents):
                if int(np.prod(sub[c, 1], dtype=np.int64)) == 0:
                    continue
                fs[rank * ncomponents + c, 0] = sub[c, 0] + np.asarray(frm)
                fs[rank * ncomponents + c, 1] = sub[c, 1]
    return fs


def basic_partitioning_ext(dim, procs, nprocs=-1, replicate=False, ext_power=None):
    """dist.h:3477-3509 (halo-extended partition)."""
    n = len(dim)
    dim = [int(d) for d in dim]
    procs = [int(d) for d in procs]
    ext = [0] * n if ext_power is None else [int(e) for e in ext_power]
    if any(e < 0 for e in ext):
        raise RuntimeError("Unsupported value for `power`")
    vol_procs = int(np.prod(procs, dtype=np.int64)) if n else 1
    P = vol_procs if nprocs < 0 else nprocs
    fs = np.zeros((P, 2, n), dtype=np.int32)
    stride = get_strides(procs, SlowToFast)
    for rank in range(vol_procs):
        cproc = _index2coor(rank, procs, stride)
        for i in range(n):
            s = min(dim[i] // procs[i] + (1 if dim[i] % procs[i] > cproc[i] else 0) + ext[i] * 2,
                    dim[i])
            fs[rank, 1, i] = s
            fs[rank, 0, i] = 0 if s == dim[i] else \
                (dim[i] // procs[i] * cproc[i] + min(cproc[i], dim[i] % procs[i]) - ext[i] + dim[i]) % dim[i]
    if replicate and vol_procs == 1:
        fs[:] = fs[0]
    return fs


def make_hole(frm, size, hole_from, hole_size, dim):
    """dist.h:3802-3825: the set (frm,size) minus (hole_from,hole_size) as a list of boxes.
    The oracle returns a *set-equivalent* description; tests compare the covered element sets,
    not the box lists."""
    n = len(dim)
    if n == 0:
        return []
    if int(np.prod(hole_size, dtype=np.int64)) == 0:
        return [(list(frm), list(size))]
    out = []
    # brute force: element membership, then emit single-element boxes is too verbose; use the
    # reference's decomposition pattern (hole | antihole | full) intersected with the range.
    parts = []
    for i in range(n):
        nf, ns = [0] * n, [0] * n
        for j in range(i):
            nf[j], ns[j] = int(hole_from[j]), int(hole_size[j])
        nf[i] = (int(hole_from[i]) + int(hole_size[i])) % int(dim[i])
        ns[i] = int(dim[i]) - int(hole_size[i])
        for j in range(i + 1, n):
            nf[j], ns[j] = 0, int(dim[j])
        parts.append((nf, ns))
    for nf, ns in parts:
        for b in intersect_boxes(nf, ns, frm, size, dim):
            if int(np.prod(b[1], dtype=np.int64)) > 0:
                out.append(b)
    return out


def _intersect_1d(f0, s0, f1, s1, d):
    """All pieces of the intersection of two periodic 1-D intervals (dist.h:353-423)."""
    if s0 == d and s1 == d:
        return [(f0, s0)]
    if s1 == d:
        return [(f0, s0)]
    if s0 == d:
        return [(f1, s1)]
    res = []

    def plain(a0, l0, a1, l1):
        fr = a0 + min(max(a1 - a0, 0), l0)
        sr = a0 + min(max(a1 + l1 - a0, 0), l0) - fr
        return fr % d, sr

    for a0, a1 in ((f0, f1), (f0, f1 + d), (f0 + d, f1)):
        fr, sr = plain(a0, s0, a1, s1)
        if sr > 0:
            res.append((fr, sr))
    return res


def intersect_boxes(f0, s0, f1, s1, dim):
    """dist.h:441-468: list of boxes of the intersection of two periodic boxes."""
    per_dim = [_intersect_1d(int(a), int(b), int(c), int(e), int(d))
               for a, b, c, e, d in zip(f0, s0, f1, s1, dim)]
    if any(len(x) == 0 for x in per_dim):
        return []
    out = [([], [])]
    for pieces in per_dim:
        out = [(f + [pf], s + [ps]) for (f, s) in out for (pf, ps) in pieces]
    return out


def box_elements(boxes, dim):
    """Set of linear indices covered by a list of periodic boxes (with multiplicity check)."""
    dim = [int(d) for d in dim]
    seen = []
    for f, s in boxes:
        if int(np.prod(s, dtype=np.int64)) == 0:
            continue
        c = (_local_coords(s, FastToSlow) + np.asarray(f, dtype=np.int64)) % np.asarray(dim)
        seen.append((c * np.asarray(get_strides(dim, FastToSlow), dtype=np.int64)).sum(axis=1))
    return np.sort(np.concatenate(seen)) if seen else np.zeros(0, dtype=np.int64)
