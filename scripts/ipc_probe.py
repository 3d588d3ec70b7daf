"""Does legacy CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) work between the ranks of this
box?  Rank 0 exports a buffer, rank 1 maps it and writes into it with a peer memcpy."""
import ctypes
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
rt = ctypes.CDLL("libcudart.so.12")
n = 1 << 20
class Handle(ctypes.Structure):  # cudaIpcMemHandle_t, passed BY VALUE to cudaIpcOpenMemHandle
    _fields_ = [("reserved", ctypes.c_ubyte * 64)]


handle = Handle()
ptr = ctypes.c_void_p()
if rank == 0:
    assert rt.cudaMalloc(ctypes.byref(ptr), n) == 0
    assert rt.cudaMemset(ptr, 0, n) == 0
    rc = rt.cudaIpcGetMemHandle(ctypes.byref(handle), ptr)
    print("rank0 cudaIpcGetMemHandle rc", rc, flush=True)
obj = [bytes(handle)]
dist.broadcast_object_list(obj, 0)
if rank == 1:
    h = Handle.from_buffer_copy(obj[0])
    mapped = ctypes.c_void_p()
    rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), Handle, ctypes.c_uint]
    rc = rt.cudaIpcOpenMemHandle(ctypes.byref(mapped), h, 1)
    print("rank1 cudaIpcOpenMemHandle rc", rc, hex(mapped.value or 0), flush=True)
    src = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    rc = rt.cudaMemcpy(mapped, ctypes.c_void_p(src.data_ptr()), n, 3)
    print("rank1 cudaMemcpy to peer rc", rc, flush=True)
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    host = (ctypes.c_ubyte * 16)()
    rt.cudaMemcpy(host, ptr, 16, 2)
    print("rank0 sees", list(host)[:4], flush=True)
dist.barrier()
