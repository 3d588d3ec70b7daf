"""Accuracy and time of the tcgen05 complex-float contraction on BASELINE config 2 (32^3 x 64, n = m = 64)
as a function of the accumulation-chain length (SBB_TC_PROMOTE, SBB_TC_KSPLIT; read once per process).
  python scripts/tc_accuracy.py [exact]     # "exact": inputs pre-truncated to TF32 (lo = 0)
Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superbblas_b200 as sb  # noqa: E402

L, LT, NV = 32, 64, 64
K = 3 * L ** 3
dev = torch.device("cuda", 0)
gpu = sb.createGpuContext(0)
g = torch.Generator(device=dev).manual_seed(7)
exact = len(sys.argv) > 1 and sys.argv[1] == "exact"


def mk():
    x = torch.empty(K * LT * NV, 2, device=dev, dtype=torch.float32)
    x.uniform_(-1.0, 1.0, generator=g)
    if exact:
        x = (x.view(torch.int32) & -8192).view(torch.float32)
    return torch.view_as_complex(x)


a, b = mk(), mk()
dimv, dimr = [3, L, L, L, LT, NV], [LT, NV, NV]
pv = np.array([[[0] * 6, dimv]], dtype=np.int32)
pr = np.array([[[0] * 3, dimr]], dtype=np.int32)
r = torch.zeros(LT * NV * NV, device=dev, dtype=torch.complex64)


def step():
    sb.contraction(1, pv, [0] * 6, dimv, dimv, 1, "cxyztn", True, [a], gpu, pv, [0] * 6, dimv, dimv, 1,
                   "cxyztm", False, [b], gpu, 0, pr, [0] * 3, dimr, dimr, 1, "tnm", [r], gpu, sb.FastToSlow)


for _ in range(3):
    step()
sb.sync(gpu)
sb.profile_enable(True)
sb.profile_read("contract_tc")
for _ in range(10):
    step()
sb.sync(gpu)
ms, n = sb.profile_read("contract_tc")
sb.profile_enable(False)
errs = []
for t in (0, 17, 63):
    A0 = a.view(NV, LT, K)[:, t, :].to(torch.complex128)
    B0 = b.view(NV, LT, K)[:, t, :].to(torch.complex128)
    ref = B0 @ A0.conj().T
    got = r.view(NV, NV, LT)[:, :, t].to(torch.complex128)
    errs.append(float((torch.linalg.norm(got - ref) / torch.linalg.norm(ref)).item()))
print(json.dumps({"promote": os.environ.get("SBB_TC_PROMOTE", "default"),
                  "ksplit": os.environ.get("SBB_TC_KSPLIT", "default"), "tf32_exact_inputs": exact,
                  "kernel_ms": ms / max(n, 1), "launches": n,
                  "TFLOP/s": 8.0 * LT * NV * NV * K / (ms / max(n, 1) * 1e-3) / 1e12 if n else None,
                  "rel_err": max(errs)}))
