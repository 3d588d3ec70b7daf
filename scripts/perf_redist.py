"""Redistribution t-slabs -> (z,t) blocks (BASELINE configs[2]) under torch.distributed.run:
per-GPU 32^3 x 64 x (4,3) x 16 complex float blocks (3.2 GB), timing with CUDA events, max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import superbblas_b200 as sb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(sb.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
comm = sb.comm_create(bytes(uid.cpu().numpy().tobytes()), world, rank, local)
gpu = sb.createGpuContext(local)
stream = torch.cuda.ExternalStream(sb.get_stream(local), device=dev)
pz = 2 if world % 2 == 0 else 1
pt = world // pz
dim = [32, 32, 32, 64, 4, 3, 16 * world]
pa = sb.basic_partitioning("xyztscn", dim, [1, 1, 1, world, 1, 1, 1], "t", world, 1)
pb = sb.basic_partitioning("xyztscn", dim, [1, 1, pz, pt, 1, 1, 1], "zt", world, 1)
nl = int(np.prod(pa[rank, 1]))
x = torch.view_as_complex(torch.rand(nl, 2, device=dev, dtype=torch.float32))
y = torch.zeros(int(np.prod(pb[rank, 1])), device=dev, dtype=torch.complex64)


def go():
    sb.copy(1, pa, 1, "xyztscn", [0] * 7, dim, dim, [x], None, gpu, pb, 1, "xyztscn", [0] * 7, dim, [y],
            None, gpu, sb.FastToSlow, sb.Copy, comm=comm)


def timed(fn, steps=10):
    import time
    for _ in range(3):
        fn()
    sb.sync(gpu)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_us = (time.perf_counter() - h0) / steps * 1e6  # host time to queue one call
        e1.record()
    sb.sync(gpu)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), host_us


out = {"world": world,
       "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_") or k.startswith("SBB_")}}
ms, host_us = timed(go)
out["redistribute_t_to_zt_c64"] = {"ms": ms, "GB/s_per_gpu": 2 * nl * 8 / ms / 1e6, "host_us_per_call": host_us}
del x, y
# periodic +1 shifts of a z,t-distributed spin-colour field (BASELINE configs[4]): 64 x 64 x 32 x 32 x (4,3) per GPU
dimf = [64, 64, 32 * pz, 32 * pt, 4, 3]
part = sb.basic_partitioning("xyztsc", dimf, [1, 1, pz, pt, 1, 1], "zt", world, 1)
nf = int(np.prod(part[rank, 1]))
for es, tag, real in (() if os.environ.get("SBB_ONLY_REDIST") else ((8, "c64", torch.float32), (16, "c128", torch.float64))):
    fx = torch.view_as_complex(torch.rand(nf, 2, device=dev, dtype=real))
    fy = torch.zeros_like(fx)
    for mu, lab in enumerate("xyzt"):
        shift = [0] * 6
        shift[mu] = 1
        ms, host_us = timed(lambda: sb.copy(1, part, 1, "xyztsc", [0] * 6, dimf, dimf, [fx], None, gpu, part, 1,
                                            "xyztsc", shift, dimf, [fy], None, gpu, sb.FastToSlow, sb.Copy,
                                            comm=comm))
        out["shift_%s_%s" % (lab, tag)] = {"ms": ms, "GB/s_per_gpu": 2 * nf * es / ms / 1e6,
                                           "host_us_per_call": host_us}
    del fx, fy
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
