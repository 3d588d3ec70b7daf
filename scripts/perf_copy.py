"""Reshuffle throughput of assorted geometries: 20 back-to-back calls between two CUDA events on the
library's stream (the tensors exceed L2)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import superbblas_b200 as sb

gpu = sb.createGpuContext(0)
stream = torch.cuda.ExternalStream(sb.get_stream(0))
out = {}


def run(name, o0, dim0, o1, dtype, from1=None, reps=20, masked=False):
    n = len(dim0)
    dim1 = [dim0[o0.index(l)] for l in o1]
    vol = int(np.prod(dim0))
    if dtype.is_complex:
        real = torch.float64 if dtype == torch.complex128 else torch.float32
        x = torch.view_as_complex(torch.rand(vol, 2, device="cuda", dtype=real))
    else:
        x = torch.rand(vol, device="cuda", dtype=dtype)
    y = torch.zeros_like(x)
    m0 = m1 = None
    if masked:  # random 50 % mask on the source, carried to the destination layout
        m0 = (torch.rand(vol, device="cuda") < 0.5).to(torch.float32)
        m1 = torch.zeros_like(m0)
        sb.copy(1, p0 := np.array([[[0] * n, dim0]], dtype=np.int32), 1, o0, [0] * n, dim0, dim0, [m0],
                None, gpu, np.array([[[0] * n, dim1]], dtype=np.int32), 1, o1, from1 or [0] * n, dim1,
                [m1], None, gpu, sb.FastToSlow, sb.Copy)
        m0, m1 = [m0], [m1]
    p0 = np.array([[[0] * n, dim0]], dtype=np.int32)
    p1 = np.array([[[0] * n, dim1]], dtype=np.int32)

    def go():
        sb.copy(1, p0, 1, o0, [0] * n, dim0, dim0, [x], m0, gpu, p1, 1, o1, from1 or [0] * n, dim1,
                [y], m1, gpu, sb.FastToSlow, sb.Copy)
    for _ in range(3):
        go()
    sb.sync(gpu)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(reps):
                go()
            e1.record()
        sb.sync(gpu)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    nbytes = 2 * vol * x.element_size()
    out[name] = {"GB/s": round(nbytes / best / 1e6, 1), "us": round(best * 1e3, 1)}
    del x, y


C128, C64 = torch.complex128, torch.complex64
run("plain_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "xyztsc", C128)
run("perm_cstzyx_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", C128)
run("perm_cstzyx_c64", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", C64)
run("perm_cstzyx_f32", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", torch.float32)
run("perm_cstzyx_f64", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", torch.float64)
run("masked_perm_cstzyx_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "cstzyx", C128, masked=True)
run("masked_plain_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "xyztsc", C128, masked=True)
run("perm_tscxyz_c128", "xyztsc", [32, 32, 32, 64, 4, 3], "tscxyz", C128)
run("perm_scxyzt_to_xyztsc_c64", "scxyzt", [4, 3, 32, 32, 32, 64], "xyztsc", C64)
run("perm_tnsxyzc_c128", "xyztscn", [16, 16, 16, 32, 4, 3, 16], "tnsxyzc", C128)
run("shift_x_c64", "xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", C64, [1, 0, 0, 0, 0, 0])
run("shift_y_c64", "xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", C64, [0, 1, 0, 0, 0, 0])
run("shift_t_c64", "xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", C64, [0, 0, 0, 1, 0, 0])
run("shift_x_c128", "xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", C128, [1, 0, 0, 0, 0, 0])
run("shift_y_c128", "xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", C128, [0, 1, 0, 0, 0, 0])
run("shift_t_c128", "xyztsc", [64, 64, 32, 32, 4, 3], "xyztsc", C128, [0, 0, 0, 1, 0, 0])
xx = torch.empty(1 << 27, device="cuda", dtype=torch.float32)
yy = torch.empty_like(xx)
for _ in range(3):
    yy.copy_(xx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    yy.copy_(xx)
e1.record()
torch.cuda.synchronize()
out["torch_copy_512MB"] = {"GB/s": round(2 * xx.numel() * 4 / (e0.elapsed_time(e1) / 20) / 1e6, 1)}
print(json.dumps(out, indent=1))
