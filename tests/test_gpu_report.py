"""The reference's SB_TRACK_TIME report (performance.h:357-441) on the GPU: per public call the
device time, the calls, and flops / bytes in the reference's units (tensor.h:1087-1088, :1593)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import superbblas_b200 as sb


def test_timings_of_copies_requests_and_contractions():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    gpu = sb.createGpuContext(0)
    dim = [8, 8, 8, 16, 4, 3]
    dim1 = dim[::-1]
    n = int(np.prod(dim))
    one = lambda d: np.array([[[0] * len(d), list(d)]], dtype=np.int32)  # noqa: E731
    x = torch.randn(n, dtype=torch.complex128, device="cuda")
    y = torch.zeros_like(x)
    r = torch.zeros(int(np.prod(dim[:4])), dtype=torch.complex128, device="cuda")
    live_before = sb.liveAllocations()
    sb.trackTime(True)
    sb.resetTimings()
    try:
        for _ in range(5):
            sb.copy(1, one(dim), 1, "xyztsc", [0] * 6, dim, dim, [x], None, gpu, one(dim1), 1, "cstzyx",
                    [0] * 6, dim1, [y], None, gpu, sb.FastToSlow, sb.Copy)
        q = sb.copy(1, one(dim), 1, "xyztsc", [0] * 6, dim, dim, [x], None, gpu, one(dim1), 1, "cstzyx",
                    [0] * 6, dim1, [y.to(torch.complex64)], None, gpu, sb.FastToSlow, sb.Copy, request=True)
        q.wait()
        for _ in range(2):
            sb.contraction(1, one(dim), [0] * 6, dim, dim, 1, "xyztsc", True, [x], gpu, one(dim), [0] * 6, dim,
                           dim, 1, "xyztsc", False, [x], gpu, 0, one(dim[:4]), [0] * 4, dim[:4], dim[:4], 1,
                           "xyzt", [r], gpu, sb.FastToSlow)
        sb.sync(gpu)
        t = sb.timings()
        assert set(t) == {"copy", "copy_begin", "wait", "contraction"}
        assert t["copy"]["calls"] == 5 and t["copy"]["bytes"] == 5 * n * (16 + 16) and t["copy"]["flops"] == 0
        assert t["copy_begin"]["calls"] == 1 and t["wait"]["calls"] == 1
        assert t["copy_begin"]["bytes"] + t["wait"]["bytes"] == n * (16 + 8)
        assert t["contraction"]["calls"] == 2
        assert t["contraction"]["flops"] == 2 * 8.0 * 8192 * 12          # 8 T M N K, m = n = 1, k = 12
        # device time was measured (events on the library stream) and is not absurd
        for name in ("copy", "contraction"):
            assert 0 < t[name]["gpu_time"] < 1.0, t
        text = sb.reportTimings()
        assert text.splitlines()[0] == "Timing of superbblas kernels:"
        assert "workspace pool, device 0" in sb.reportCacheUsage() or "copy plans" in sb.reportCacheUsage()
        assert sb.liveAllocations() == live_before      # the calls returned every workspace they took
    finally:
        sb.trackTime(False)
        sb.resetTimings()
