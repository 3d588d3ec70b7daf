"""The C-ABI library loads without a GPU and exports every symbol include/superbblas_b200.h
declares; host-only entry points (partitions, make_hole) agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import superbblas_b200 as sb
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "superbblas_b200.h")).read()
    names = set(re.findall(r"\b(sb[bk]_[a-z_0-9]+)\s*\(", header))
    assert len(names) >= 20
    lib = ctypes.CDLL(sb.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert b"sm_100a" in ctypes.cast(lib.sbb_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_context_layout_matches_reference():
    # class Context {enum platform plat; int device;} (platform.h:757-765)
    assert ctypes.sizeof(sb.Context) == 8
    assert sb.Context.plat.offset == 0 and sb.Context.device.offset == 4
    assert sb.createGpuContext(3).plat == 1 and sb.createGpuContext(3).device == 3
    assert sb.createCpuContext().plat == 0


def test_partitions_match_oracle():
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(1, 7))
        order = "".join(rng.permutation(list("xyztscn"))[:n])
        dim = [int(rng.integers(1, 13)) for _ in range(n)]
        labels = "".join(rng.permutation(list(order))[:int(rng.integers(1, n + 1))])
        nprocs = int(rng.integers(1, 25))
        procs = sb.partitioning_distributed_procs(order, dim, labels, nprocs)
        assert procs == O.partitioning_distributed_procs(order, dim, labels, nprocs)
        ncomp = int(rng.integers(1, 4))
        P = int(np.prod(procs))
        assert np.array_equal(sb.basic_partitioning(order, dim, procs, labels, P, ncomp),
                              O.basic_partitioning(order, dim, procs, labels, P, ncomp))
        ext = [int(rng.integers(0, 3)) for _ in range(n)]
        assert np.array_equal(sb.basic_partitioning(dim, procs, P, False, ext),
                              O.basic_partitioning_ext(dim, procs, P, False, ext))


def test_known_partition_answers():
    # tests/dist.cpp:103-125 of the reference
    assert sb.partitioning_distributed_procs("xyzt", [8, 8, 8, 16], "zt", 8) == [1, 1, 2, 4]
    p = sb.basic_partitioning("xyztscn", [32, 32, 32, 64, 4, 3, 128], [1, 1, 2, 4, 1, 1, 1], "zt", 8)
    # rank = 4*jz + jt
    for jz in range(2):
        for jt in range(4):
            assert list(p[4 * jz + jt, 0]) == [0, 0, 16 * jz, 16 * jt, 0, 0, 0]
            assert list(p[4 * jz + jt, 1]) == [32, 32, 16, 16, 4, 3, 128]


def test_make_hole_matches_oracle_sets():
    rng = np.random.default_rng(6)
    for _ in range(200):
        n = int(rng.integers(1, 4))
        dim = [int(rng.integers(1, 7)) for _ in range(n)]
        frm = [int(rng.integers(0, d)) for d in dim]
        size = [int(rng.integers(1, d + 1)) for d in dim]
        hf = [int(rng.integers(0, d)) for d in dim]
        hs = [int(rng.integers(0, d + 1)) for d in dim]
        boxes = sb.make_hole(frm, size, hf, hs, dim)
        a = O.box_elements(boxes, dim)
        b = O.box_elements(O.make_hole(frm, size, hf, hs, dim), dim)
        assert np.array_equal(a, b)
        assert len(np.unique(a)) == len(a)  # boxes do not overlap


def test_no_cuda_device_fails_loudly():
    if sb.getGpuDevicesCount() > 0:
        return
    p = np.array([[[0], [4]]], dtype=np.int32)
    x, y = np.zeros(4), np.zeros(4)
    try:
        sb.copy(1, p, 1, "x", [0], [4], [4], [x], None, sb.createCpuContext(), p, 1, "x", [0], [4],
                [y], None, sb.createCpuContext(), sb.FastToSlow, sb.Copy)
    except RuntimeError as e:
        assert "CUDA" in str(e) or "device" in str(e)
    else:
        raise AssertionError("copy must not succeed without a GPU (no CPU compute path)")


def test_cxx_front_end_host_checks(tmp_path):
    """include/superbblas.h compiled with plain g++ (no CUDA toolkit needed by callers): partition
    generators, make_hole and the detail:: range helpers used by the reference's tests/dist.cpp,
    against brute-force enumeration (tests/cxx/host_api_test.cpp).  Runs without a GPU."""
    import shutil
    import subprocess
    import superbblas_b200
    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(superbblas_b200.LIB_PATH)
    exe = str(tmp_path / "host_api_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"),
                        os.path.join(root, "tests", "cxx", "host_api_test.cpp"), "-o", exe,
                        "-L" + libdir, "-lsuperbblas_b200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host api ok" in r.stdout, r.stdout + r.stderr
    assert "Timing of superbblas kernels" not in r.stdout
    # the reference's reports, switched on the reference's way
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, SB_TRACK_TIME="1", SB_TRACK_MEMORY="1"))
    assert r.returncode == 0 and "host api ok" in r.stdout, r.stdout + r.stderr
    assert "Timing of superbblas kernels" in r.stdout and "copy : " in r.stdout


def test_reports_of_the_public_calls():
    """sbb_report / sbb_track_time (performance.h:357-518): names, counts and text format; the
    numbers themselves need a GPU (tests/test_gpu_report.py)."""
    sb.trackTime(False)
    sb.resetTimings()
    assert sb.reportTimings() == ""
    sb.trackTime(True)
    try:
        p = np.array([[[0, 0], [4, 4]]], dtype=np.int32)
        x, y = np.zeros(16), np.zeros(16)
        cpu = sb.createCpuContext()
        for _ in range(3):
            try:
                sb.copy(1, p, 1, "ab", [0, 0], [4, 4], [4, 4], [x], None, cpu, p, 1, "ba", [0, 0], [4, 4],
                        [y], None, cpu, sb.FastToSlow, sb.Copy)
            except RuntimeError:
                assert sb.getGpuDevicesCount() == 0   # fails only without a device; counted all the same
        t = sb.timings()
        assert set(t) == {"copy"} and t["copy"]["calls"] == 3
        if sb.getGpuDevicesCount() > 0:
            assert t["copy"]["bytes"] == 3 * 16 * (8 + 8)
        lines = sb.reportTimings().splitlines()
        assert lines[0] == "Timing of superbblas kernels:" and lines[2].startswith("copy : ")
        for word in ("gpu_time:", "calls:", "flops:", "bytes:", "GFLOPs_single:", "GBYTES/s:", "intensity:"):
            assert word in lines[2]
        # a buffer that is too small is reported with the size needed
        lib = ctypes.CDLL(sb.LIB_PATH)
        small = ctypes.create_string_buffer(8)
        needed = ctypes.c_size_t(0)
        assert lib.sbb_report(0, small, ctypes.c_size_t(8), ctypes.byref(needed)) == 2
        assert needed.value == len(sb.reportTimings()) + 1
        assert lib.sbb_report(7, small, ctypes.c_size_t(8), ctypes.byref(needed)) == 1
        assert "copy plans" in sb.reportCacheUsage()
        sb.resetTimings()
        assert sb.timings() == {}
    finally:
        sb.trackTime(False)


@pytest.mark.parametrize("program", ["dist.cpp", "contract.cpp"])
def test_reference_tests_compile_against_the_drop_in_header(program):
    """The reference's own test programs, unchanged, are valid callers of include/superbblas.h --
    also in their MPI configuration (SUPERBBLAS_USE_MPI, compile-checked against the stand-in
    tests/cxx/mpi_stub/mpi.h).  Syntax and template instantiation only; the GPU tests run them."""
    import shutil
    import subprocess
    src = os.path.join("/root/reference/tests", program)
    if not os.path.exists(src) or not shutil.which("g++"):
        pytest.skip("needs /root/reference and g++")
    for defs in (["-DSUPERBBLAS_USE_GPU"], ["-DSUPERBBLAS_USE_GPU", "-DSUPERBBLAS_USE_MPI"]):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w"] + defs +
                           ["-I" + os.path.join(ROOT, "tests", "cxx", "mpi_stub"),
                            "-I" + os.path.join(ROOT, "include"), src], capture_output=True, text=True)
        assert r.returncode == 0, (defs, r.stderr[-2000:])


def test_front_end_marshalling_of_coordinates():
    """superbblas_b200.api builds the int arrays of the C ABI from whatever sequence the caller has
    and remembers them by VALUE: an array modified in place between two calls is a new value."""
    from superbblas_b200 import api
    assert list(api._iv([1, 2, 3], 3)) == [1, 2, 3]
    assert list(api._iv((np.int64(4), np.int32(5)))) == [4, 5]
    assert list(api._iv(np.array([[1, 2], [3, 4]], dtype=np.int64))) == [1, 2, 3, 4]
    assert list(api._iv([[1, 2], [3, 4]])) == [1, 2, 3, 4]
    assert list(api._iv([np.array([1, 2]), np.array([3, 4])])) == [1, 2, 3, 4]
    a = np.array([7, 8, 9], dtype=np.int32)
    first = api._iv(a)
    assert api._iv(a) is first                       # same value: same marshalled array
    a[1] = 80
    assert list(api._iv(a)) == [7, 80, 9] and list(first) == [7, 8, 9]
    v = np.arange(24, dtype=np.int32).reshape(2, 3, 4)[:, :, ::2]   # non-contiguous view
    assert list(api._iv(v)) == [int(x) for x in v.reshape(-1)]
    with pytest.raises(RuntimeError, match="wrong length"):
        api._iv([1, 2], 3)
    with pytest.raises(RuntimeError, match="wtf"):
        api._partition(np.zeros((2, 2, 3), dtype=np.int32), 3, 3)
    # the same partition through copy_plan as nested lists, int64 and int32 arrays
    dim = [4, 6]
    p32 = sb.basic_partitioning("ab", dim, [2, 1], "a", 2, 1)
    q32 = sb.basic_partitioning("ab", dim, [1, 2], "b", 2, 1)
    plans = [sb.copy_plan(8, p, 1, "ab", [0, 0], dim, dim, q, 1, "ab", [0, 1], dim, 2, 0, sb.FastToSlow, sb.Copy)
             for p, q in ((p32, q32), (p32.astype(np.int64), q32.astype(np.int64)), (p32.tolist(), q32.tolist()))]
    assert len(plans[0][0]) > 0
    assert plans[0] == plans[1] == plans[2]
