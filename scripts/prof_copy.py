"""Runs the reshuffle cases once each (after one warm-up) so that ncu can capture the permute
kernel of every geometry:  ncu --set full -k regex:permute_kernel ... python scripts/prof_copy.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import superbblas_b200 as sb

gpu = sb.createGpuContext(0)


def case(o0, dim0, o1, dtype, from1=None, reps=2):
    n = len(dim0)
    dim1 = [dim0[o0.index(l)] for l in o1]
    vol = int(np.prod(dim0))
    x = torch.view_as_complex(torch.rand(vol, 2, device="cuda",
                                         dtype=torch.float64 if dtype == torch.complex128 else torch.float32))
    y = torch.zeros_like(x)
    p0 = np.array([[[0] * n, dim0]], dtype=np.int32)
    p1 = np.array([[[0] * n, dim1]], dtype=np.int32)
    for _ in range(reps):
        sb.copy(1, p0, 1, o0, [0] * n, dim0, dim0, [x], None, gpu, p1, 1, o1, from1 or [0] * n, dim1,
                [y], None, gpu, sb.FastToSlow, sb.Copy)
    sb.sync(gpu)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "perm128"):
    case("xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", torch.complex128)
if which in ("all", "perm64"):
    case("xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", torch.complex64)
if which in ("all", "plain"):
    case("xyztsc", [32, 32, 32, 64, 4, 3], "xyztsc", torch.complex128)
if which in ("all", "shiftx"):
    case("xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", torch.complex64, [1, 0, 0, 0, 0, 0])
if which in ("all", "shiftt"):
    case("xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", torch.complex64, [0, 0, 0, 1, 0, 0])
if which in ("masked",):
    # even-site masks on both sides of the xyztsc -> cstzyx permutation (one pass: source mask read with the data)
    dim0 = [32, 32, 32, 64, 4, 3]
    dim1 = dim0[::-1]
    vol = int(np.prod(dim0))
    x = torch.view_as_complex(torch.rand(vol, 2, device="cuda", dtype=torch.float64))
    y = torch.zeros_like(x)
    idx = torch.arange(vol, device="cuda")
    par = (idx % 32 + (idx // 32) % 32 + (idx // 1024) % 32 + (idx // 32768) % 64) % 2
    m0 = (par == 0).to(torch.float32)
    m1 = m0.view(3, 4, 64, 32, 32, 32).permute(5, 4, 3, 2, 1, 0).contiguous().view(-1)
    p0 = np.array([[[0] * 6, dim0]], dtype=np.int32)
    p1 = np.array([[[0] * 6, dim1]], dtype=np.int32)
    for _ in range(2):
        sb.copy(1, p0, 1, "xyztsc", [0] * 6, dim0, dim0, [x], [m0], gpu, p1, 1, "cstzyx", [0] * 6, dim1, [y],
                [m1], gpu, sb.FastToSlow, sb.Copy)
    sb.sync(gpu)
print("done")
