"""What does NCCL itself deliver for the redistribution's message pattern (each rank sends 2 x 1.2 GB
and receives 2 x 1.2 GB)?  torch.distributed batch_isend_irecv and all_to_all_single."""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1200 * (1 << 20)
peers_out = [(rank + 1) % world, (rank + 2) % world][: max(1, min(2, world - 1))]
peers_in = [(rank - 1) % world, (rank - 2) % world][: max(1, min(2, world - 1))]
sb = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in peers_out]
rb = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in peers_in]


def p2p():
    ops = [dist.P2POp(dist.isend, s, p) for s, p in zip(sb, peers_out)] + \
          [dist.P2POp(dist.irecv, r, p) for r, p in zip(rb, peers_in)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms = timeit(p2p)
out = {"world": world, "p2p_ms": ms, "p2p_out_GBs_per_gpu": len(peers_out) * n / ms / 1e6}
a = torch.empty(world * (n // 2), dtype=torch.uint8, device="cuda")
b = torch.empty_like(a)
ms = timeit(lambda: dist.all_to_all_single(b, a))
out.update(a2a_ms=ms, a2a_out_GBs_per_gpu=(world - 1) * (n // 2) / ms / 1e6)
# one-directional peer copy through torch (cudaMemcpyPeer) for reference: not possible across processes
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
