"""The FP32-from-TF32 operand split of the tcgen05 contraction kernel (kernels_contract_tc.cu,
`tf32_hi` / `split_store`), restated in numpy with the same integer operations and checked for the
properties DESIGN.md §4.2b relies on: hi is x rounded to the nearest TF32 value, x - hi is exact,
and hi*hi + hi*lo + lo*hi reproduces the FP32 product to ~2^-21.  (The tensor core's own
accumulation error is a measured property of the hardware: profiles/r2_tc_accuracy_sweep.jsonl.)"""
import numpy as np


def tf32_hi(x):
    """(bits + 0x1000) & 0xffffe000, as in the kernel"""
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def samples(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n) * np.exp2(rng.integers(-30, 30, n))
    return x.astype(np.float32)


def test_hi_is_round_to_nearest_tf32():
    x = samples(200000, 1)
    hi = tf32_hi(x)
    assert not np.any(hi.view(np.uint32) & np.uint32(0x1FFF))          # 10 explicit mantissa bits
    # nearest: the error is at most half a TF32 ulp (2^-11 relative to the binade of x)
    m, e = np.frexp(x.astype(np.float64))
    ulp = np.exp2(e - 11.0)                                                 # TF32 spacing in x's binade
    assert np.all(np.abs(hi.astype(np.float64) - x) <= ulp / 2)
    # ties round away from zero, like cvt.rna.tf32.f32
    tie = np.array([1.0 + 2.0 ** -11, -(1.0 + 2.0 ** -11), 1.0 + 3 * 2.0 ** -11], dtype=np.float32)
    assert np.array_equal(tf32_hi(tie), np.array([1.0 + 2.0 ** -10, -(1.0 + 2.0 ** -10), 1.0 + 2.0 ** -9],
                                                 dtype=np.float32))
    # values that are TF32 already, zeros, infinities pass through
    exact = np.array([0.0, -0.0, 1.0, 1.5, -3.25, np.inf, -np.inf, 2.0 ** -126], dtype=np.float32)
    assert np.array_equal(tf32_hi(exact).view(np.uint32), exact.view(np.uint32))


def test_lo_is_exact_and_small():
    x = samples(200000, 2)
    hi = tf32_hi(x)
    d = x - hi                                                              # float32 subtraction, as in the kernel
    assert np.array_equal(d.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))
    lo = tf32_hi(d)
    # what the split drops: at most half a TF32 ulp of lo, i.e. 2^-22 relative to x
    rest = x.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64)
    assert np.all(np.abs(rest) <= np.abs(x.astype(np.float64)) * 2.0 ** -21)
    assert np.all(np.abs(lo.astype(np.float64)) <= np.abs(x.astype(np.float64)) * 2.0 ** -10)


def test_three_products_give_fp32_accuracy():
    # complex dot products of config-2 length, the four real blocks as the kernel forms them
    # (A' = [Re; Im], B' = [Re; Im]); accumulation in double here: this test is about the split only
    k = 98304
    rng = np.random.default_rng(7)
    a = (rng.uniform(-1, 1, k) + 1j * rng.uniform(-1, 1, k)).astype(np.complex64)
    b = (rng.uniform(-1, 1, k) + 1j * rng.uniform(-1, 1, k)).astype(np.complex64)

    def parts(v):
        hi = tf32_hi(v)
        return hi.astype(np.float64), tf32_hi(v - hi).astype(np.float64)

    def dot3(u, v):
        uh, ul = parts(u)
        vh, vl = parts(v)
        return np.sum(uh * vh) + (np.sum(uh * vl) + np.sum(ul * vh))

    rr, ii = dot3(a.real.copy(), b.real.copy()), dot3(a.imag.copy(), b.imag.copy())
    ri, ir = dot3(a.real.copy(), b.imag.copy()), dot3(a.imag.copy(), b.real.copy())
    got = complex(rr + ii, ri - ir)                                         # conj(a) . b
    want = np.vdot(a.astype(np.complex128), b.astype(np.complex128))
    scale = np.sum(np.abs(a.astype(np.complex128)) * np.abs(b.astype(np.complex128)))
    assert abs(got - want) <= 2.0 ** -20 * scale
    # the hi.hi term alone (plain TF32) is two orders of magnitude worse on the same data
    single = (np.sum(parts(a.real.copy())[0] * parts(b.real.copy())[0]) +
              np.sum(parts(a.imag.copy())[0] * parts(b.imag.copy())[0]))
    assert abs(single - want.real) > 100 * abs(got.real - want.real)
