"""Helpers for the GPU parity tests: move numpy components to torch CUDA tensors and back."""
import numpy as np
import torch

import superbblas_b200 as sb


def to_dev(arrs, device="cuda:0"):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in arrs]


def to_host(tensors):
    return [t.cpu().numpy() for t in tensors]


def run_copy(case, v0, v1, host0=(), host1=(), mask0=None, mask1=None):
    """sb.copy with every part a component of one process on cuda:0 (parts listed in host0/host1
    stay in host memory, i.e. CPU contexts; their masks too)."""
    P0, P1 = case["p0"].shape[0], case["p1"].shape[0]
    gpu, cpu = sb.createGpuContext(0), sb.createCpuContext()
    d0 = [np.ascontiguousarray(a.copy()) if i in host0 else to_dev([a])[0] for i, a in enumerate(v0)]
    d1 = [np.ascontiguousarray(a.copy()) if j in host1 else to_dev([a])[0] for j, a in enumerate(v1)]
    ctx0 = [cpu if i in host0 else gpu for i in range(P0)]
    ctx1 = [cpu if j in host1 else gpu for j in range(P1)]
    m0 = m1 = None
    if mask0 is not None:
        m0 = [np.ascontiguousarray(a.copy()) if i in host0 else to_dev([a])[0]
              for i, a in enumerate(mask0)]
    if mask1 is not None:
        m1 = [np.ascontiguousarray(a.copy()) if j in host1 else to_dev([a])[0]
              for j, a in enumerate(mask1)]
    sb.copy(case["alpha"], case["p0"], P0, case["o0"], case["from0"], case["size0"], case["dim0"],
            d0, m0, ctx0, case["p1"], P1, case["o1"], case["from1"], case["dim1"], d1, m1, ctx1,
            case["co"], case["copyadd"])
    sb.sync(gpu)
    return [x if isinstance(x, np.ndarray) else x.cpu().numpy() for x in d1]


def run_contraction(case, v0, v1, vr):
    gpu = sb.createGpuContext(0)
    d0, d1, dr = to_dev(v0), to_dev(v1), to_dev(vr)
    sb.contraction(case["alpha"], case["p0"], case["from0"], case["size0"], case["dim0"], len(v0),
                   case["o0"], case["conj0"], d0, gpu, case["p1"], case["from1"], case["size1"],
                   case["dim1"], len(v1), case["o1"], case["conj1"], d1, gpu, case["beta"],
                   case["pr"], case["fromr"], case["sizer"], case["dimr"], len(vr), case["o_r"], dr,
                   gpu, case["co"])
    sb.sync(gpu)
    return to_host(dr)
