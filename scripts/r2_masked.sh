#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_copy.py -m gpu -x -q > gpurun_out/r2_pytest_d.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_d.log)"
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
import superbblas_b200 as sb
gpu = sb.createGpuContext(0)
stream = torch.cuda.ExternalStream(sb.get_stream(0))
dim0 = [32, 32, 32, 64, 4, 3]; dim1 = dim0[::-1]; vol = int(np.prod(dim0))
p0 = np.array([[[0] * 6, dim0]], dtype=np.int32); p1 = np.array([[[0] * 6, dim1]], dtype=np.int32)
idx = torch.arange(vol, device="cuda")
par = (idx % 32 + (idx // 32) % 32 + (idx // 1024) % 32 + (idx // 32768) % 64) % 2
m0 = (par == 0).to(torch.float32)
m1 = m0.view(3, 4, 64, 32, 32, 32).permute(5, 4, 3, 2, 1, 0).contiguous().view(-1)
rnd = (torch.rand(vol, device="cuda") < 0.5).to(torch.float32)
rnd1 = rnd.view(3, 4, 64, 32, 32, 32).permute(5, 4, 3, 2, 1, 0).contiguous().view(-1)
for name, cdt, rdt in (("c128", torch.complex128, torch.float64), ("c64", torch.complex64, torch.float32)):
    x = torch.view_as_complex(torch.rand(vol, 2, device="cuda", dtype=rdt)); y = torch.zeros_like(x)
    for tag, o1, d1, pp1, a, b in (("even cstzyx", "cstzyx", dim1, p1, m0, m1), ("random cstzyx", "cstzyx", dim1, p1, rnd, rnd1),
                                   ("even same-order", "xyztsc", dim0, p0, m0, m0), ("unmasked cstzyx", "cstzyx", dim1, p1, None, None)):
        fn = lambda: sb.copy(1, p0, 1, "xyztsc", [0] * 6, dim0, dim0, [x], None if a is None else [a], gpu, pp1, 1, o1, [0] * 6, d1, [y],
                             None if b is None else [b], gpu, sb.FastToSlow, sb.Copy)
        for _ in range(3): fn()
        sb.sync(gpu)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(20): fn()
            e1.record()
        sb.sync(gpu); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%-5s %-18s %.4f ms  %6.0f GB/s (reference byte count)  frac %.3f" % (name, tag, ms, 2 * vol * x.element_size() / ms / 1e6, 2 * vol * x.element_size() / ms / 1e6 / 6545.3))
PY
