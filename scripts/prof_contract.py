"""Runs BASELINE configs[1] (V^H V -> [t,n,m], 32^3 x 64, n=m=64) twice on complex float or complex
double operands so that ncu can capture contract_mma_kernel:
  ncu --set full -k regex:contract_mma -s 1 -c 1 ... python scripts/prof_contract.py c64"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import superbblas_b200 as sb

gpu = sb.createGpuContext(0)
cplx = torch.complex64 if (len(sys.argv) > 1 and sys.argv[1] == "c64") else torch.complex128
real = torch.float32 if cplx == torch.complex64 else torch.float64
dimv, dimr = [3, 32, 32, 32, 64, 64], [64, 64, 64]
pv = np.array([[[0] * 6, dimv]], dtype=np.int32)
pr = np.array([[[0] * 3, dimr]], dtype=np.int32)
n = int(np.prod(dimv))
a = torch.view_as_complex(torch.rand(n, 2, device="cuda", dtype=real))
b = torch.view_as_complex(torch.rand(n, 2, device="cuda", dtype=real))
r = torch.zeros(int(np.prod(dimr)), device="cuda", dtype=cplx)
for _ in range(2):
    sb.contraction(1, pv, [0] * 6, dimv, dimv, 1, "cxyztn", True, [a], gpu, pv, [0] * 6, dimv, dimv, 1,
                   "cxyztm", False, [b], gpu, 0, pr, [0] * 3, dimr, dimr, 1, "tnm", [r], gpu, sb.FastToSlow)
sb.sync(gpu)
print("done")
