"""Which contraction kernel a problem takes, and how it is split, examined on the host
(`sbk_contract_describe`; without a device the B200's 148 SMs are assumed).  Pins the dispatch rules of
DESIGN.md §4.2-4.2c for the BASELINE shapes and for the skinny shapes of the reference's tests/dist.cpp
(the reference's own heuristics for those: blas.h:686-699)."""
import os

import pytest

import superbblas_b200 as sb


def distillation(T, n, K, conj0=True):
    """cxyztn^H . cxyztm -> tnm after the planner merged c,x,y,z: K contiguous in both operands"""
    return sb.contract_desc(T=[(T, K, K, 1)], M=[(n, K * T, 0, T)], N=[(n, 0, K * T, T * n)],
                            K=[(K, 1, 1, 0)], conj0=conj0)


def fields(text):
    w = text.split()
    return {x.split("=")[0]: x.split("=")[1] for x in w if "=" in x}


@pytest.fixture(autouse=True)
def no_forced_kernel(monkeypatch):
    monkeypatch.delenv("SBB_CONTRACT_KERNEL", raising=False)
    monkeypatch.delenv("SBB_TC_KSPLIT", raising=False)
    monkeypatch.delenv("SBB_TC_PROMOTE", raising=False)


def test_headline_contraction_fills_whole_waves():
    # BASELINE configs[1] in complex double: FP64 tensor-pipe kernel, 64 x 64 tiles, 2 CTAs per SM;
    # 64 tiles x 37 K-slices = 2368 CTAs = exactly 8 waves of 148 x 2
    text = sb.contract_describe(distillation(64, 64, 3 * 32 ** 3), sb.C128)
    f = fields(text)
    assert text.startswith("mma f64 tile=64x64x8") and f["loader"] == "cp.async"
    assert (f["ksplit"], f["ctas"]) == ("37", "2368") and 2368 % (148 * 2) == 0
    assert f["a_kfast"] == f["b_kfast"] == "1"
    # configs[3] per GPU on z2 x t4: 24 time slices, 2 x 2 tiles of the 128 x 128 result
    f = fields(sb.contract_describe(distillation(24, 128, 3 * 48 * 48 * 24), sb.C128))
    assert int(f["ctas"]) == 24 * 4 * int(f["ksplit"]) and int(f["ctas"]) % (148 * 2) == 0
    # real double and float operands share the kernel (float: widened when the fragments are read)
    assert sb.contract_describe(distillation(64, 64, 3 * 32 ** 3), sb.F64).startswith("mma f64 tile")
    assert "(float operands)" in sb.contract_describe(distillation(64, 64, 3 * 32 ** 3), sb.F32)


def test_complex_float_takes_the_tcgen05_kernel_when_eligible():
    text = sb.contract_describe(distillation(64, 64, 3 * 32 ** 3), sb.C64)
    f = fields(text)
    assert text.startswith("tcgen05 tf32x3") and f["loader"] == "tma" and f["promote"] == "4"
    assert int(f["ctas"]) == 64 * int(f["ksplit"]) and int(f["smem"]) <= 227 * 1024
    # one CTA per SM: the split leaves at most 3 % of the last wave empty
    waves = -(-int(f["ctas"]) // 148)
    assert int(f["ctas"]) / (waves * 148) >= 0.97
    # short contractions, tiny tiles and strided K stay on the FP64-pipe kernel
    assert sb.contract_describe(distillation(64, 64, 128), sb.C64).startswith("mma f64")
    strided = sb.contract_desc(T=[(8, 64 * 4096, 64 * 4096, 1)], M=[(64, 1, 0, 8)], N=[(64, 0, 1, 512)],
                               K=[(4096, 64, 64, 0)])
    assert sb.contract_describe(strided, sb.C64).startswith("mma f64")
    os.environ["SBB_CONTRACT_KERNEL"] = "mma"
    try:
        assert sb.contract_describe(distillation(64, 64, 3 * 32 ** 3), sb.C64).startswith("mma f64")
    finally:
        del os.environ["SBB_CONTRACT_KERNEL"]


@pytest.mark.parametrize("m,n,k,kernel", [
    (1, 1, 49152, "dot"), (4, 4, 49152, "dot"), (12, 12, 49152, "dot"),
    (49152, 1, 1, "row"), (49152, 3, 3, "row"), (49152, 12, 12, "row"), (49152, 16, 16, "row"),
])
def test_skinny_shapes_of_the_reference_dist_program(m, n, k, kernel):
    # dist.cpp's batched GEMMs in column-major BLAS layout: v0[m,k,b] v1[k,n,b] -> vr[m,n,b], batch 32
    b = 32
    d = sb.contract_desc(T=[(b, m * k, k * n, m * n)], M=[(m, 1, 0, 1)], N=[(n, 0, k, m)], K=[(k, m, 1, 0)])
    for dt in (sb.C64, sb.C128, sb.F32, sb.F64):
        assert sb.contract_describe(d, dt).split()[0] == kernel, (m, n, k, dt)


def test_site_wise_colour_spin_contraction_of_config_1():
    # BASELINE configs[0]: xyztsc^H . xyztsc -> xyzt: 8192 sites, K = 12, m = n = 1
    d = sb.contract_desc(T=[(8192, 1, 1, 1)], K=[(12, 8192, 8192, 0)], conj0=True)
    assert sb.contract_describe(d, sb.C128) == "row rows=8192 small=1 K=12 big_operand=v0"
    # ... and the contract.cpp-style variant with 4 x 4 free labels: xyztscN^H . xyztscn -> xyztNn
    d = sb.contract_desc(T=[(8192, 1, 1, 1)], M=[(4, 8192 * 12, 0, 8192)], N=[(4, 0, 8192 * 12, 4 * 8192)],
                         K=[(12, 8192, 8192, 0)], conj0=True)
    assert sb.contract_describe(d, sb.C128).split()[0] in ("row", "simt")


def test_degenerate_problems():
    assert sb.contract_describe(sb.contract_desc(T=[(0, 1, 1, 1)], K=[(4, 1, 1, 0)]), sb.F64) == "empty"
    with pytest.raises(RuntimeError, match="unsupported type"):
        sb.contract_describe(distillation(2, 2, 8), sb.I32)
