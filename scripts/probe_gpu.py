"""One-off measurements on the B200 box: FP64 GEMM peaks (the roofline denominator that
MEASURED_PEAKS.json lacks) and first timings of the copy / contraction kernels."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import superbblas_b200 as sb


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    out = {}
    N = 8192
    a = torch.randn(N, N, device="cuda", dtype=torch.float64)
    b = torch.randn(N, N, device="cuda", dtype=torch.float64)
    ms = timeit(lambda: torch.matmul(a, b), n=5)
    out["dgemm_8192_tflops"] = 2 * N ** 3 / ms / 1e9
    N = 4096
    a = torch.randn(N, N, device="cuda", dtype=torch.complex128)
    b = torch.randn(N, N, device="cuda", dtype=torch.complex128)
    ms = timeit(lambda: torch.matmul(a, b), n=5)
    out["zgemm_4096_tflops"] = 8 * N ** 3 / ms / 1e9
    del a, b
    # cuBLAS on the distillation shape: batched (64 x K)^H (K x 64)
    Lt, nv, K = 64, 64, 3 * 32 ** 3
    A = torch.randn(nv, Lt, K, 2, device="cuda", dtype=torch.float64)
    B = torch.randn(nv, Lt, K, 2, device="cuda", dtype=torch.float64)
    Ac, Bc = torch.view_as_complex(A), torch.view_as_complex(B)
    ms = timeit(lambda: torch.einsum("ntk,mtk->tnm", Ac.conj(), Bc), n=3, warm=1)
    out["torch_einsum_config2_ms"] = ms
    out["torch_einsum_config2_tflops"] = 8 * Lt * nv * nv * K / ms / 1e9

    gpu = sb.createGpuContext(0)
    stream = torch.cuda.ExternalStream(sb.get_stream(0))
    dimv, dimr = [3, 32, 32, 32, Lt, nv], [Lt, nv, nv]
    pv = np.array([[[0] * 6, dimv]], dtype=np.int32)
    pr = np.array([[[0] * 3, dimr]], dtype=np.int32)
    r = torch.zeros(Lt * nv * nv, device="cuda", dtype=torch.complex128)
    a1, b1 = Ac.view(-1), Bc.view(-1)

    def contr():
        sb.contraction(1, pv, [0] * 6, dimv, dimv, 1, "cxyztn", True, [a1], gpu, pv, [0] * 6, dimv,
                       dimv, 1, "cxyztm", False, [b1], gpu, 0, pr, [0] * 3, dimr, dimr, 1, "tnm",
                       [r], gpu, sb.FastToSlow)
    with torch.cuda.stream(stream):
        ms = timeit(contr, n=5, warm=2)
    out["sbb_contraction_config2_ms"] = ms
    out["sbb_contraction_config2_tflops"] = 8 * Lt * nv * nv * K / ms / 1e9
    ref = torch.einsum("ntk,mtk->mnt", Ac.conj(), Bc).contiguous().view(-1)
    out["sbb_contraction_config2_relerr"] = (torch.linalg.norm(r - ref) / torch.linalg.norm(ref)).item()
    del A, B, Ac, Bc, a1, b1, ref

    # copies
    def copy_case(name, o0, dim0, o1, dtype, from1=None):
        n = len(dim0)
        dim1 = [dim0[o0.index(l)] for l in o1]
        vol = int(np.prod(dim0))
        x = torch.randn(vol * (2 if dtype.is_complex else 1), device="cuda",
                        dtype=torch.float64 if dtype in (torch.complex128, torch.float64) else torch.float32)
        x = torch.view_as_complex(x.view(-1, 2)) if dtype.is_complex else x
        y = torch.zeros_like(x)
        p0 = np.array([[[0] * n, dim0]], dtype=np.int32)
        p1 = np.array([[[0] * n, dim1]], dtype=np.int32)
        f1 = from1 or [0] * n

        def go():
            sb.copy(1, p0, 1, o0, [0] * n, dim0, dim0, [x], None, gpu, p1, 1, o1, f1, dim1, [y],
                    None, gpu, sb.FastToSlow, sb.Copy)
        with torch.cuda.stream(stream):
            ms = timeit(go, n=10, warm=3)
        nbytes = 2 * vol * x.element_size()
        out["copy_%s_GBs" % name] = nbytes / ms / 1e6
        out["copy_%s_ms" % name] = ms

    copy_case("xyztsc_cstzyx_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", torch.complex128)
    copy_case("xyztsc_cstzyx_c64", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", torch.complex64)
    copy_case("xyztsc_tscxyz_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "tscxyz", torch.complex128)
    copy_case("plain_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "xyztsc", torch.complex128)
    copy_case("shift_x_c64", "xyztsc", [64, 64, 64, 16, 4, 3], "xyztsc", torch.complex64, [1, 0, 0, 0, 0, 0])
    copy_case("shift_t_c64", "xyztsc", [64, 64, 64, 16, 4, 3], "xyztsc", torch.complex64, [0, 0, 0, 1, 0, 0])
    copy_case("scxyzt_xyztsc_c64", "scxyzt", [4, 3, 32, 32, 32, 64], "xyztsc", torch.complex64)
    xx = torch.empty(1 << 28, device="cuda", dtype=torch.float32)
    yy = torch.empty_like(xx)
    ms = timeit(lambda: yy.copy_(xx), n=10)
    out["torch_copy_GBs"] = 2 * xx.numel() * 4 / ms / 1e6
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
