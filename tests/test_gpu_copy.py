"""GPU parity: superbblas_b200.copy (CUDA kernels through the C ABI) against the oracle, bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import superbblas_b200 as sb
from oracle import oracle as O
from tests import cases as C


@pytest.fixture(scope="module")
def gu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_util
    return gpu_util


@pytest.mark.parametrize("seed", range(10))
def test_random_copies_bit_exact(gu, seed):
    rng = np.random.default_rng(500 + seed)
    for it in range(40):
        case = C.random_copy_case(rng, max_dim=7)
        v0, v1 = C.make_copy_data(case, seed * 100 + it, consistent=case["copyadd"] == 0)
        want = C.oracle_copy(case, v0, v1)
        got = gu.run_copy(case, v0, v1)
        for j, (g, w) in enumerate(zip(got, want)):
            assert C.bits_equal(g, w), (seed, it, j, case)


@pytest.mark.parametrize("seed", range(6))
def test_masked_copies_bit_exact(gu, seed):
    """Masks (tensor.h:1022-1027, dist.h:944-970): compatible mask pairs as the
    reference requires them, on random geometries, partitions, types, alpha, Copy/Add; some
    components (and their masks) in host memory."""
    rng = np.random.default_rng(1500 + seed)
    for it in range(40):
        case = C.random_copy_case(rng, max_dim=7)
        v0, v1 = C.make_copy_data(case, seed * 100 + it, consistent=True)
        m0, m1 = C.make_masks(case, seed * 100 + it, density=[0.5, 0.1, 0.9][it % 3])
        want = C.oracle_copy(case, v0, v1, m0, m1)
        h0 = [i for i in range(len(v0)) if seed % 2 and rng.random() < 0.3]
        h1 = [j for j in range(len(v1)) if seed % 2 and rng.random() < 0.3]
        got = gu.run_copy(case, v0, v1, host0=h0, host1=h1, mask0=m0, mask1=m1)
        for j, (g, w) in enumerate(zip(got, want)):
            assert C.bits_equal(g, w), (seed, it, j, case)


def test_one_sided_masks(gu):
    """Only one of the two masks given: an element moves iff the given mask is nonzero at its
    source (mask0) or at its destination (mask1); everything else is untouched."""
    rng = np.random.default_rng(1600)
    for it in range(60):
        case = C.random_copy_case(rng, max_dim=6)
        v0, v1 = C.make_copy_data(case, 700 + it, consistent=True)
        m0, m1 = C.make_masks(case, 700 + it)
        if it % 2:
            m0 = None
        else:
            m1 = None
        want = C.oracle_copy(case, v0, v1, m0, m1)
        got = gu.run_copy(case, v0, v1, mask0=m0, mask1=m1)
        for j, (g, w) in enumerate(zip(got, want)):
            assert C.bits_equal(g, w), (it, j, case)


def test_masked_even_odd_lattice(gu):
    """Chroma-style even/odd checkerboard on a 16^3 x 32 x 4 x 3 field, permuted "xyztsc" ->
    "cstzyx": the even copy followed by the odd copy equals the unmasked copy and each half leaves
    the other half untouched."""
    import torch
    dim0, dim1 = [16, 16, 16, 32, 4, 3], [3, 4, 32, 16, 16, 16]
    p0, p1 = np.array([[[0] * 6, dim0]], dtype=np.int32), np.array([[[0] * 6, dim1]], dtype=np.int32)
    vol = int(np.prod(dim0))
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.view_as_complex(torch.randn(vol, 2, generator=g, device="cuda", dtype=torch.float64))
    # parity of x+y+z+t; FastToSlow: x fastest => torch shape is the reversed dim list
    ax = [torch.arange(d, device="cuda") for d in dim0]
    par0 = (ax[0].view(1, 1, 1, 1, 1, -1) + ax[1].view(1, 1, 1, 1, -1, 1) + ax[2].view(1, 1, 1, -1, 1, 1) +
            ax[3].view(1, 1, -1, 1, 1, 1) + 0 * ax[4].view(1, -1, 1, 1, 1, 1) +
            0 * ax[5].view(-1, 1, 1, 1, 1, 1)) % 2
    even0 = (par0 == 0).to(torch.float32).contiguous().view(-1)
    odd0 = 1 - even0
    perm = lambda m: m.view(*reversed(dim0)).permute(5, 4, 3, 2, 1, 0).contiguous().view(-1)
    even1, odd1 = perm(even0), perm(odd0)
    gpu = sb.createGpuContext(0)

    def run(m0, m1, dst):
        sb.copy(1, p0, 1, "xyztsc", [0] * 6, dim0, dim0, [a], m0, gpu, p1, 1, "cstzyx", [0] * 6, dim1,
                [dst], m1, gpu, sb.FastToSlow, sb.Copy)
        sb.sync(gpu)
    full = torch.zeros_like(a)
    run(None, None, full)
    sentinel = torch.full_like(a, 7 - 3j)
    b = sentinel.clone()
    run([even0], [even1], b)
    assert torch.equal(torch.view_as_real(b)[odd1 != 0], torch.view_as_real(sentinel)[odd1 != 0])
    assert torch.equal(torch.view_as_real(b)[even1 != 0], torch.view_as_real(full)[even1 != 0])
    run([odd0], [odd1], b)
    assert torch.equal(torch.view_as_real(b), torch.view_as_real(full))


def test_host_components_are_staged_through_the_gpu(gu):
    rng = np.random.default_rng(600)
    for it in range(40):
        case = C.random_copy_case(rng)
        v0, v1 = C.make_copy_data(case, 900 + it, consistent=case["copyadd"] == 0)
        want = C.oracle_copy(case, v0, v1)
        h0 = [i for i in range(len(v0)) if rng.random() < 0.5]
        h1 = [j for j in range(len(v1)) if rng.random() < 0.5]
        got = gu.run_copy(case, v0, v1, host0=h0, host1=h1)
        for j, (g, w) in enumerate(zip(got, want)):
            assert C.bits_equal(g, w), (it, j, case)


@pytest.mark.parametrize("co", [0, 1])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64, np.float64, np.float32, np.int32])
def test_config1_permutation(gu, co, dtype):
    """BASELINE config 1: "xyztsc" -> "cstzyx" on 8^3 x 16 x 4 x 3."""
    dim0, dim1 = [8, 8, 8, 16, 4, 3], [3, 4, 16, 8, 8, 8]
    p0, p1 = np.array([[[0] * 6, dim0]], dtype=np.int32), np.array([[[0] * 6, dim1]], dtype=np.int32)
    case = dict(alpha=1, p0=p0, o0="xyztsc", from0=[0] * 6, size0=dim0, dim0=dim0, p1=p1,
                o1="cstzyx", from1=[0] * 6, dim1=dim1, co=co, copyadd=0, T=np.dtype(dtype),
                Q=np.dtype(dtype))
    v0, v1 = C.make_copy_data(case, 1)
    want = C.oracle_copy(case, v0, v1)
    got = gu.run_copy(case, v0, v1)
    assert C.bits_equal(got[0], want[0])


def _strides(dims):
    s, acc = [], 1
    for d in dims:
        s.append(acc)
        acc *= d
    return s


@pytest.mark.parametrize("es_dtype", [np.int32, np.complex64, np.complex128])
def test_kernel_level_transposes(gu, es_dtype):
    """sbk_permute_copy on boxes that exercise every variant: direct, tiled, ragged tiles, many
    dims, sub-boxes with offsets, promoted element widths."""
    import torch
    rng = np.random.default_rng(700)
    shapes = [[33, 65], [64, 64], [3, 4, 17, 9], [2, 3, 2, 3, 2, 3, 2, 3, 2, 3], [1000, 3],
              [3, 1000], [5, 7, 11, 13], [128, 2, 128], [12, 32, 32], [4096], [1], [2, 2],
              [31, 1, 29, 2]]
    for shape in shapes:
        n = len(shape)
        for trial in range(3):
            perm = list(rng.permutation(n))
            # the source is a sub-box of a bigger tensor half of the time
            pad = [int(rng.integers(0, 3)) for _ in shape] if trial else [0] * n
            sdims = [s + p for s, p in zip(shape, pad)]
            sfrom = [int(rng.integers(0, p + 1)) for p in pad]
            ss = _strides(sdims)
            ddims = [shape[k] for k in perm]
            ds_perm = _strides(ddims)
            dstride = [0] * n
            for pos, k in enumerate(perm):
                dstride[k] = ds_perm[pos]
            src = C.fill(int(np.prod(sdims)), es_dtype, 42)
            dst0 = C.fill(int(np.prod(ddims)), es_dtype, 43)
            soff = sum(f * s for f, s in zip(sfrom, ss))
            want = dst0.copy()
            idx_s = np.array([soff])
            idx_d = np.array([0])
            for k in range(n):
                idx_s = (idx_s[:, None] + np.arange(shape[k]) * ss[k]).reshape(-1)
                idx_d = (idx_d[:, None] + np.arange(shape[k]) * dstride[k]).reshape(-1)
            want[idx_d] = src[idx_s]
            ts, td = torch.from_numpy(src).cuda(), torch.from_numpy(dst0).cuda()
            sb.permute_copy(sb.box_desc(shape, ss, dstride, soff, 0), ts, td)
            sb.sync(sb.createGpuContext(0))
            assert C.bits_equal(td.cpu().numpy(), want), (shape, perm, pad)


def test_large_permutation_round_trip(gu):
    """Size-independent property at a BASELINE-like size: permuting and permuting back is the
    identity, and the permuted tensor matches a torch permute (32^3 x 16 x 12 complex double)."""
    import torch
    dim0 = [32, 32, 32, 16, 4, 3]
    o0, o1 = "xyztsc", "cstzyx"
    dim1 = [3, 4, 16, 32, 32, 32]
    vol = int(np.prod(dim0))
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(vol, 2, generator=g, device="cuda", dtype=torch.float64)
    a = torch.view_as_complex(a)
    b = torch.zeros_like(a)
    c = torch.zeros_like(a)
    gpu = sb.createGpuContext(0)
    p0 = np.array([[[0] * 6, dim0]], dtype=np.int32)
    p1 = np.array([[[0] * 6, dim1]], dtype=np.int32)
    sb.copy(1, p0, 1, o0, [0] * 6, dim0, dim0, [a], None, gpu, p1, 1, o1, [0] * 6, dim1, [b], None,
            gpu, sb.FastToSlow, sb.Copy)
    sb.copy(1, p1, 1, o1, [0] * 6, dim1, dim1, [b], None, gpu, p0, 1, o0, [0] * 6, dim0, [c], None,
            gpu, sb.FastToSlow, sb.Copy)
    sb.sync(gpu)
    assert torch.equal(torch.view_as_real(a), torch.view_as_real(c))
    # FastToSlow: first label fastest => torch shape is the reversed dim list
    ref = a.view(*reversed(dim0)).permute(5, 4, 3, 2, 1, 0).contiguous().view(-1)
    assert torch.equal(torch.view_as_real(ref), torch.view_as_real(b))


def test_multi_component_redistribution_and_shift(gu):
    """BASELINE configs 3 and 5 with 8 'ranks' emulated as 8 components on one GPU (the reference's
    own trick): t-slabs -> (z,t) blocks, periodic shifts, halo fill."""
    P = 8
    dim = [8, 8, 8, 16, 4, 3, 4]
    p0 = sb.basic_partitioning("xyztscn", dim, [1, 1, 1, P, 1, 1, 1], "t", P, 1)
    p1 = sb.basic_partitioning("xyztscn", dim, [1, 1, 2, P // 2, 1, 1, 1], "zt", P, 1)
    case = dict(alpha=1, p0=p0, o0="xyztscn", from0=[0] * 7, size0=dim, dim0=dim, p1=p1,
                o1="xyztscn", from1=[0] * 7, dim1=dim, co=1, copyadd=0, T=np.dtype(np.complex64),
                Q=np.dtype(np.complex64))
    v0, v1 = C.make_copy_data(case, 1)
    assert all(C.bits_equal(g, w) for g, w in zip(gu.run_copy(case, v0, v1), C.oracle_copy(case, v0, v1)))

    dim = [8, 8, 8, 16, 4, 3]
    p = sb.basic_partitioning("xyztsc", dim, [1, 1, 2, 4, 1, 1], "zt", P, 1)
    for dtype in (np.complex64, np.complex128):
        for shift in ([1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0],
                      [0, 0, 7, 15, 0, 0]):
            case = dict(alpha=1, p0=p, o0="xyztsc", from0=[0] * 6, size0=dim, dim0=dim, p1=p,
                        o1="xyztsc", from1=shift, dim1=dim, co=1, copyadd=0, T=np.dtype(dtype),
                        Q=np.dtype(dtype))
            v0, v1 = C.make_copy_data(case, 2)
            assert all(C.bits_equal(g, w)
                       for g, w in zip(gu.run_copy(case, v0, v1), C.oracle_copy(case, v0, v1)))
    # halo fill (dist.cpp:459-504)
    ph = sb.basic_partitioning(dim, [1, 1, 2, 4, 1, 1], P, False, [1, 1, 1, 1, 0, 0])
    pc = sb.basic_partitioning(dim, [1, 1, 2, 4, 1, 1], P)
    case = dict(alpha=1, p0=pc, o0="xyztsc", from0=[0] * 6, size0=dim, dim0=dim, p1=ph, o1="xyztsc",
                from1=[0] * 6, dim1=dim, co=1, copyadd=0, T=np.dtype(np.complex64),
                Q=np.dtype(np.complex64))
    v0, v1 = C.make_copy_data(case, 3)
    assert all(C.bits_equal(g, w) for g, w in zip(gu.run_copy(case, v0, v1), C.oracle_copy(case, v0, v1)))


def test_golden_vectors_on_gpu(gu):
    """Outputs of the real reference (tests/golden/*.npz) reproduced by the CUDA path."""
    from tests.test_golden import load_golden
    for name, g in load_golden():
        if g["kind"] not in ("copy", "masked_copy"):
            continue
        got = gu.run_copy(g["case"], g["v0"], g["v1"], mask0=g["m0"], mask1=g["m1"])
        for j, (x, w) in enumerate(zip(got, g["want"])):
            assert C.bits_equal(x, w), (name, j)


def test_local_copy_wrapper(gu):
    """local_copy (north-star API; signature of the reference's tests/local.cpp:97): one component on
    each side, with and without masks, against the oracle's copy with a single PartitionItem."""
    import torch
    rng = np.random.default_rng(1700)
    gpu = sb.createGpuContext(0)
    for it in range(30):
        case = C.random_copy_case(rng, nparts0=1, nparts1=1, max_dim=6)
        case["p0"] = np.array([[[0] * len(case["dim0"]), case["dim0"]]], dtype=np.int32)
        case["p1"] = np.array([[[0] * len(case["dim1"]), case["dim1"]]], dtype=np.int32)
        v0, v1 = C.make_copy_data(case, 800 + it, consistent=True)
        m0, m1 = C.make_masks(case, 800 + it) if it % 2 else (None, None)
        want = C.oracle_copy(case, v0, v1, m0, m1)
        a, b = torch.from_numpy(v0[0]).cuda(), torch.from_numpy(v1[0]).cuda()
        dm0 = torch.from_numpy(m0[0]).cuda() if m0 is not None else None
        dm1 = torch.from_numpy(m1[0]).cuda() if m1 is not None else None
        sb.local_copy(case["alpha"], case["o0"], case["from0"], case["size0"], case["dim0"], a, dm0,
                      gpu, case["o1"], case["from1"], case["dim1"], b, dm1, gpu, case["co"],
                      case["copyadd"])
        sb.sync(gpu)
        assert C.bits_equal(b.cpu().numpy(), want[0]), (it, case)


def test_errors(gu):
    import torch
    gpu = sb.createGpuContext(0)
    p = np.array([[[0, 0], [2, 2]]], dtype=np.int32)
    x = torch.zeros(4, device="cuda", dtype=torch.float64)
    y = torch.zeros(4, device="cuda", dtype=torch.float64)
    with pytest.raises(RuntimeError, match="Invalid copy operation"):
        sb.copy(1, p, 1, "xy", [0, 0], [2, 2], [2, 2], [x], None, gpu, p, 1, "xz", [0, 0], [2, 2],
                [y], None, gpu, sb.FastToSlow, sb.Copy)
    with pytest.raises(RuntimeError):
        sb.copy(1, p, 1, "xyz", [0, 0], [2, 2], [2, 2], [x], None, gpu, p, 1, "xy", [0, 0], [2, 2],
                [y], None, gpu, sb.FastToSlow, sb.Copy)
    with pytest.raises(RuntimeError, match="masks must be float32"):
        sb.copy(1, p, 1, "xy", [0, 0], [2, 2], [2, 2], [x], [x], gpu, p, 1, "xy", [0, 0], [2, 2],
                [y], None, gpu, sb.FastToSlow, sb.Copy)
    # empty and degenerate inputs
    p0 = np.array([[[0, 0], [0, 0]]], dtype=np.int32)
    e = torch.zeros(0, device="cuda", dtype=torch.float64)
    sb.copy(1, p0, 1, "xy", [0, 0], [0, 0], [2, 2], [e], None, gpu, p, 1, "xy", [0, 0], [2, 2], [y],
            None, gpu, sb.FastToSlow, sb.Copy)
    sb.sync(gpu)


def test_storage_staging_pattern(gu):
    """The hot-path part of the reference's S3T save / load (storage.h:1029 local_save, :1149
    local_load): a sub-box of a GPU tensor is permuted, scaled and converted (double -> float) on
    the fly into a contiguous HOST buffer in the file's label order, and read back from the host
    buffer into a sub-box of a GPU tensor.  Both directions against the oracle, bit for bit; the
    save side also through a Request (the copy-back to the host completes at wait)."""
    import torch
    rng = np.random.default_rng(77)
    gpu, cpu = sb.createGpuContext(0), sb.createCpuContext()
    dim0 = [6, 5, 4, 3]                       # tensor "xyzn" on the GPU
    from0, size0 = [1, 0, 2, 0], [4, 5, 2, 3]  # the sub-box that is saved (wraps in z)
    o_file = "nzxy"
    size1 = [size0["xyzn".index(c)] for c in o_file]
    one = lambda d: np.array([[[0] * len(d), list(d)]], dtype=np.int32)  # noqa: E731
    save = dict(alpha=0.5, p0=one(dim0), o0="xyzn", from0=from0, size0=size0, dim0=dim0, p1=one(size1),
                o1=o_file, from1=[0] * 4, dim1=size1, co=0, copyadd=0, T=np.dtype(np.complex128),
                Q=np.dtype(np.complex64))
    v0, v1 = C.make_copy_data(save, 1)
    want = C.oracle_copy(save, v0, v1)
    src = torch.from_numpy(v0[0]).cuda()
    host = v1[0].copy()
    q = sb.copy(0.5, save["p0"], 1, "xyzn", from0, size0, dim0, [src], None, gpu, save["p1"], 1, o_file,
                [0] * 4, size1, [host], None, cpu, sb.SlowToFast, sb.Copy, request=True)
    q.wait()
    assert C.bits_equal(host, want[0])
    # load: the host buffer (file order) into a sub-box of another GPU tensor, converting back
    load = dict(alpha=1, p0=one(size1), o0=o_file, from0=[0] * 4, size0=size1, dim0=size1, p1=one(dim0),
                o1="xyzn", from1=from0, dim1=dim0, co=0, copyadd=0, T=np.dtype(np.complex64),
                Q=np.dtype(np.complex128))
    w0, w1 = C.make_copy_data(load, 2)
    w0 = [host.copy()]
    wantl = C.oracle_copy(load, w0, w1)
    dst = torch.from_numpy(w1[0].copy()).cuda()
    sb.copy(1, load["p0"], 1, o_file, [0] * 4, size1, size1, [w0[0]], None, cpu, load["p1"], 1, "xyzn", from0,
            dim0, [dst], None, gpu, sb.SlowToFast, sb.Copy)
    sb.sync(gpu)
    assert C.bits_equal(dst.cpu().numpy(), wantl[0])
