#!/bin/bash
# After the front-end marshalling change: the GPU tests that go through superbblas_b200.api, and the
# per-call time of the configs[0] permutation through the Python front end.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_copy.py tests/test_gpu_loopback.py tests/test_gpu_contraction.py \
    -m gpu -x -q > gpurun_out/r2_pytest_api.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2_pytest_api.log)"
timeout 60 python - > gpurun_out/r2_small_call.json 2> gpurun_out/r2_small_call.err <<'PY'
import json, time, numpy as np, torch
import superbblas_b200 as sb
dim = [8, 8, 8, 16, 4, 3]; dim1 = dim[::-1]
p0 = np.array([[[0] * 6, dim]], dtype=np.int32); p1 = np.array([[[0] * 6, dim1]], dtype=np.int32)
x = torch.randn(int(np.prod(dim)), dtype=torch.complex128, device="cuda"); y = torch.zeros_like(x)
gpu = sb.createGpuContext(0)
def call():
    sb.copy(1, p0, 1, "xyztsc", [0] * 6, dim, dim, [x], None, gpu, p1, 1, "cstzyx", [0] * 6, dim1, [y], None, gpu,
            sb.FastToSlow, sb.Copy)
for _ in range(200):
    call()
sb.sync(gpu)
t = time.perf_counter()
n = 5000
for _ in range(n):
    call()
issue = (time.perf_counter() - t) / n
sb.sync(gpu)
total = (time.perf_counter() - t) / n
print(json.dumps({"workload": "configs[0] permutation xyztsc->cstzyx c128 through the Python front end",
                  "calls": n, "host_issue_us_per_call": issue * 1e6, "us_per_call": total * 1e6}))
PY
echo "small rc=$?"; cat gpurun_out/r2_small_call.json
