#!/usr/bin/env python
"""bench.py — the tensor hot path of superbblas on B200: contraction TFLOP/s (headline) and
reshuffle GB/s, against measured rooflines, next to the reference's CPU build on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1] per GPU): the distillation contraction
    R[t,n,m] = sum_{c,x,y,z} conj(V0[c,x,y,z,t,n]) * V1[c,x,y,z,t,m],   complex double,
32^3 x 64 sites and n = m = 64 vectors PER GPU (2.0616e11 flop, 12.9 GB of operands).  With N GPUs
the lattice grows to 32 x 32 x (32*Pz) x (64*Pt), (Pz,Pt) = (1,1),(2,1),(2,2),(2,4), partitioned on
z,t like BASELINE configs[3]; the z-halves produce partial sums that are reduced across ranks into
an output partitioned on t (NCCL send/recv + add).  Weak scaling: the work per GPU is fixed.

A step = one `contraction` call through the public API with device-resident operands (`value`), or
with HOST operands in pinned memory, staged by the library (`e2e`).  Timing: CUDA events on the
library's stream, barrier + synchronize on both sides, max over ranks.  Operands (12 GiB) are far
larger than the 126 MB L2, so no flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L, LT, NV = 32, 64, 64          # per-GPU lattice extent, time extent, vectors
FLOP_PER_GPU = 8.0 * LT * NV * NV * 3 * L ** 3
CPU_SAMPLE_T = 16               # time slices of the CPU sample (1/4 of one GPU's work)


def grid_for(n):
    pz = 2 if n % 2 == 0 else 1
    return pz, n // pz


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref/libsbref.so) on the host cores
# ---------------------------------------------------------------------------------------------------

def cpu_contraction_sample(reps, warmup):
    """Times the reference's CPU contraction on a bounded sample of the workload: the same
    32^3 lattice and n=m=64 vectors with CPU_SAMPLE_T of the 64 time slices."""
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")  # the reference parallelises over the batch
    from oracle import ref as R
    kind = "reference"
    if not R.available():
        raise RuntimeError("oracle/_ref/libsbref.so is missing (run `make -C oracle` where "
                           "/root/reference exists)")
    dimv, dimr = [3, L, L, L, CPU_SAMPLE_T, NV], [CPU_SAMPLE_T, NV, NV]
    pv = np.array([[[0] * 6, dimv]], dtype=np.int32)
    pr = np.array([[[0] * 3, dimr]], dtype=np.int32)
    rng = np.random.default_rng(0x5B5B0000 + 2 * 16)
    n = int(np.prod(dimv))
    a = rng.random(n) + 1j * rng.random(n)
    b = rng.random(n) + 1j * rng.random(n)
    c = np.zeros(int(np.prod(dimr)), dtype=np.complex128)
    times = []
    for i in range(warmup + reps):
        t0 = time.perf_counter()
        R.contraction(1, pv, [0] * 6, dimv, dimv, "cxyztn", True, [a], pv, [0] * 6, dimv, dimv,
                      "cxyztm", False, [b], 0, pr, [0] * 3, dimr, dimr, "tnm", [c], "FastToSlow")
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    flop = 8.0 * CPU_SAMPLE_T * NV * NV * 3 * L ** 3
    return dict(total_s=sum(times), steps=len(times), flop_per_step=flop, cores=cores, kind=kind,
                sample="same contraction with %d of 64 time slices (%.3g flop per step), "
                       "OMP_NUM_THREADS=%d OPENBLAS_NUM_THREADS=1" % (CPU_SAMPLE_T, flop, cores))


def cpu_copy_sample(reps=5):
    """The reference's CPU copy on a bounded sample of the reshuffle workload: "xyztsc" -> "cstzyx" on
    16^3 x 32 x 4 x 3 complex double (25 MB; warm calls, i.e. with its index vectors cached)."""
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    from oracle import ref as R
    dim0, dim1 = [16, 16, 16, 32, 4, 3], [3, 4, 32, 16, 16, 16]
    p0 = np.array([[[0] * 6, dim0]], dtype=np.int32)
    p1 = np.array([[[0] * 6, dim1]], dtype=np.int32)
    n = int(np.prod(dim0))
    rng = np.random.default_rng(1)
    a = rng.random(n) + 1j * rng.random(n)
    b = np.zeros(n, dtype=np.complex128)
    best = None
    for i in range(reps + 2):
        t0 = time.perf_counter()
        R.copy(1, p0, "xyztsc", [0] * 6, dim0, dim0, [a], p1, "cstzyx", [0] * 6, dim1, [b], "FastToSlow", 0)
        dt = time.perf_counter() - t0
        if i >= 2:
            best = dt if best is None else min(best, dt)
    return {"GB/s": 2 * n * 16 / best / 1e9, "ms": best * 1e3, "cores": cores, "kind": "reference",
            "sample": "16^3 x 32 x 4 x 3 complex double (25 MB), warm (index vectors cached)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        r = cpu_contraction_sample(args.steps, args.warmup)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0]}))
        return 0
    value = r["flop_per_step"] * r["steps"] / r["total_s"] / 1e12
    line = {
        "impl": "reference", "metric": "contraction TFLOP/s (distillation V^H V -> [t,n,m])",
        "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": args.warmup, "ms_per_step": 1e3 * r["total_s"] / r["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def bind_host_to_gpu_numa(torch, local):
    """N > 1: keep this rank's host threads -- and therefore its pinned staging buffers (first touch)
    -- on the CPUs NVML names as local to its GPU, so that the host->device copies of the e2e leg do
    not cross sockets.  Best effort: returns the number of CPUs bound to, or None when anything is
    missing (SBB_BENCH_NUMA=0 disables it).  Not used at N = 1, where the CPU baseline needs every core."""
    if os.environ.get("SBB_BENCH_NUMA", "1") == "0":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        u = str(torch.cuda.get_device_properties(local).uuid)
        if not u.startswith("GPU-"):
            u = "GPU-" + u
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(u)
        except TypeError:
            h = pynvml.nvmlDeviceGetHandleByUUID(u.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return None


def workload_config(n):
    pz, pt = grid_for(n)
    return {"workload": "BASELINE configs[1] per GPU: contraction cxyztn^H . cxyztm -> tnm, "
                        "32^3x64 sites and n=m=64 complex double per GPU",
            "lattice": [L, L, L * pz, LT * pt], "vectors": NV, "partition": "z%d x t%d" % (pz, pt),
            "coor_order": "FastToSlow", "l2": "operands (12.9 GB per GPU) exceed L2; no flush"}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms",
                 "100", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                w = [x.strip() for x in line.split(",")]
                if len(w) < 10:
                    continue
                try:
                    sm.append(float(w[2])), mx.append(float(w[3]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if w[6 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            # under load = the upper half of the samples (the sampler also sees idle gaps)
            hi = sorted(sm)[len(sm) // 2:]
            out.update(sm_mhz=float(np.median(hi)), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the reshuffle measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import superbblas_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_affinity = bind_host_to_gpu_numa(torch, local) if world > 1 else None
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(sb.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = sb.comm_create(bytes(uid.cpu().numpy().tobytes()), world, rank, local)
    gpu = sb.createGpuContext(local)
    cpu = sb.createCpuContext()
    stream = torch.cuda.ExternalStream(sb.get_stream(local), device=dev)

    def barrier():
        sb.sync(gpu)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload -----------------------------------------------------------------------------------
    pz, pt = grid_for(world)
    dimv = [3, L, L, L * pz, LT * pt, NV]
    dimr = [LT * pt, NV, NV]
    pv = sb.basic_partitioning("cxyztn", dimv, [1, 1, 1, pz, pt, 1], "zt", world, 1)
    pr = sb.basic_partitioning("tnm", dimr, [world, 1, 1], "t", world, 1)
    nloc = int(np.prod(pv[rank, 1]))
    nout = int(np.prod(pr[rank, 1]))
    g = torch.Generator(device=dev).manual_seed(0x5B5B0000 + 2 * 16 + rank)
    mk = lambda: torch.view_as_complex(  # noqa: E731
        torch.rand(nloc, 2, generator=g, device=dev, dtype=torch.float64) * 2 - 1)
    a, b = mk(), mk()
    r = torch.zeros(nout, device=dev, dtype=torch.complex128)

    def step(va, vb, vr, ctx):
        sb.contraction(1, pv, [0] * 6, dimv, dimv, 1, "cxyztn", True, [va], ctx, pv, [0] * 6, dimv,
                       dimv, 1, "cxyztm", False, [vb], ctx, 0, pr, [0] * 3, dimr, dimr, 1, "tnm",
                       [vr], ctx, sb.FastToSlow, comm=comm)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- FP64 roofline denominator, measured live (MEASURED_PEAKS.json has no FP64 entry) ----------
    def zgemm_peak():
        n = 4096
        x = torch.randn(n, n, device=dev, dtype=torch.complex128)
        y = torch.randn(n, n, device=dev, dtype=torch.complex128)
        best = 1e9
        for i in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, y)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                best = min(best, e0.elapsed_time(e1))
        return 8.0 * n ** 3 / best / 1e9
    fp64_peak = zgemm_peak()

    # ---- timed region: device-resident operands --------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sb.profile_enable(False)
    sb.launch_count(reset=True)
    ms = timed(lambda: step(a, b, r, gpu), args.steps, max(args.warmup, 3))
    launches = sb.launch_count()
    # the warm-up launches are counted too; report the timed ones only
    launches = launches * args.steps // (args.steps + max(args.warmup, 3))
    value = FLOP_PER_GPU * world * args.steps / (ms * 1e-3) / 1e12
    # dominant kernel, timed with events on its own stream over another pass of the same steps
    sb.profile_enable(True)
    sb.profile_read("contract_mma")
    timed(lambda: step(a, b, r, gpu), args.steps, 1)
    kms, kn = sb.profile_read("contract_mma")
    sb.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = kms / max(kn, 1)
    achieved = FLOP_PER_GPU / (kernel_ms * 1e-3) / 1e12

    # result sanity: trace-like checksum against a cuBLAS evaluation of one time slice
    K = 3 * L ** 3
    A0 = a.view(NV, LT, K)[:, 0, :]
    B0 = b.view(NV, LT, K)[:, 0, :]
    check_ok = None
    if world == 1:
        ref0 = (B0 @ A0.conj().T)  # [m][n]
        got0 = r.view(NV, NV, LT)[:, :, 0]
        check_ok = bool((torch.linalg.norm(got0 - ref0) / torch.linalg.norm(ref0)).item() < 1e-12)

    # ---- the same contraction on complex float operands (N=1): products and K sum on the FP64 tensor pipe ---
    contraction_c64 = None
    if world == 1 and not args.no_extras:
        af, bf = a.to(torch.complex64), b.to(torch.complex64)
        rf = torch.zeros(nout, device=dev, dtype=torch.complex64)
        sb.profile_enable(True)
        sb.profile_read("contract_mma")
        ms_f = timed(lambda: step(af, bf, rf, gpu), 10, 3)
        kms_f, kn_f = sb.profile_read("contract_mma")
        sb.profile_enable(False)
        ref0f = (B0 @ A0.conj().T)
        got0f = rf.view(NV, NV, LT)[:, :, 0].to(torch.complex128)
        contraction_c64 = {"TFLOP/s": FLOP_PER_GPU * 10 / (ms_f * 1e-3) / 1e12, "ms": ms_f / 10,
                           "kernel_ms": kms_f / max(kn_f, 1),
                           "frac_of_fp64_tensor_peak": FLOP_PER_GPU / (kms_f / max(kn_f, 1) * 1e-3) / 1e12 / fp64_peak,
                           "rel_err_vs_c128": float((torch.linalg.norm(got0f - ref0f) /
                                                     torch.linalg.norm(ref0f)).item()),
                           "min_bytes": 2 * nloc * 8 + nout * 8}
        del af, bf, rf

    # ---- e2e: HOST operands in pinned memory through the same public call ---------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    ha = torch.empty(nloc, dtype=torch.complex128, pin_memory=True)
    hb = torch.empty(nloc, dtype=torch.complex128, pin_memory=True)
    hr = torch.zeros(nout, dtype=torch.complex128, pin_memory=True)
    ha.copy_(a), hb.copy_(b)
    torch.cuda.synchronize()
    ms_e2e = timed(lambda: step(ha, hb, hr, cpu), e2e_steps, 1)
    e2e_value = FLOP_PER_GPU * world * e2e_steps / (ms_e2e * 1e-3) / 1e12
    h2d = 2 * nloc * 16 * world
    d2h = nout * 16 * world
    if world == 1 and check_ok:
        check_ok = bool(torch.allclose(hr.to(dev), r, rtol=1e-12, atol=1e-9))
    del ha, hb

    # ---- extras: reshuffle GB/s -------------------------------------------------------------------------
    extras = {}
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        pass
    if not args.no_extras:
        del a, b
        torch.cuda.empty_cache()
        extras = reshuffle_extras(sb, torch, dist, dev, gpu, stream, comm, rank, world, timed,
                                  hbm_peak)

    # ---- CPU reference for the reshuffle (rank 0, N=1 only): same permutation, 16^3 x 32 x 4 x 3 ------------
    if world == 1 and not args.no_cpu and not args.no_extras:
        try:
            extras["cpu_reference_permute_xyztsc_cstzyx_c128"] = cpu_copy_sample()
        except Exception as e:  # noqa: BLE001
            extras["cpu_reference_permute_xyztsc_cstzyx_c128"] = {"unavailable": str(e)}

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        try:
            # in a process of its own (the reference arm of this script): inside this process two
            # OpenMP runtimes (torch's and the reference library's) fight over the cores and the
            # same sample runs 3x slower
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                                  "--steps", "3", "--warmup", "1"], capture_output=True, text=True,
                                 timeout=600).stdout
            ref = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
            if "unavailable" in ref:
                raise RuntimeError(ref["unavailable"])
            cpu_baseline = ref["cpu_baseline"]
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"value": None, "unit": "TFLOP/s", "cores": os.cpu_count(),
                            "kind": "reference", "sample": "unavailable: %s" % e}

    if rank == 0:
        line = {
            "metric": "contraction TFLOP/s (distillation V^H V -> [t,n,m])",
            "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic", "config": workload_config(world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the
                         # ncu --set full capture in profiles/r1_contract_mma_ncu.txt
                         "traffic": 12.885870e9 + 140.346368e6,
                         "kernel": "contract_mma_kernel<double2>", "kernel_ms": kernel_ms,
                         "flop_per_launch": FLOP_PER_GPU,
                         "peak_source": "FP64 path: measured live, cuBLAS ZGEMM 4096^3 via "
                                        "torch.matmul, best of 5 (MEASURED_PEAKS.json has no FP64 "
                                        "entry)"},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": launches, "clocks": clocks, "result_check": check_ok,
            "host_cpus_bound_to_gpu_numa": host_affinity,
            "reshuffle": extras, "contraction_c64": contraction_c64, "hbm_peak_gbs": hbm_peak,
            "hbm_peak_source": hbm_src,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def reshuffle_extras(sb, torch, dist, dev, gpu, stream, comm, rank, world, timed, hbm_peak):
    """Reshuffle GB/s = moved elements x (sizeof T + sizeof Q) / time (the reference's `memops`,
    tensor.h:1087): label permutation, periodic shift, and (N > 1) redistribution t -> (z,t)."""
    out = {}
    local = dev.index

    def field(nbytes_elem, n):
        dt = torch.complex64 if nbytes_elem == 8 else torch.complex128
        return torch.view_as_complex(torch.rand(n, 2, device=dev, dtype=torch.float32
                                                if nbytes_elem == 8 else torch.float64))

    def record(name, nbytes, fn, steps=20):
        ms = timed(fn, steps, 3)  # the throughput figure: no per-kernel events in the stream
        sb.profile_enable(True)   # second pass: the copy kernels alone, bracketed by events
        sb.profile_read("permute")
        timed(fn, steps, 1)
        kms, kn = sb.profile_read("permute")
        sb.profile_enable(False)
        gbs = nbytes * world * steps / (ms * 1e-3) / 1e9
        kernel_ms = kms / steps  # all permute kernels of one step (they run back to back)
        out[name] = {"GB/s": gbs, "ms": ms / steps, "per_gpu_frac_of_hbm": gbs / world / hbm_peak,
                     "kernel_ms": kernel_ms, "kernel_launches_per_step": kn / steps,
                     "kernel_GB/s_per_gpu": nbytes / kernel_ms / 1e6 if kn else None}

    # (a) label permutation "xyztsc" -> "cstzyx", 32^3 x 64 x 4 x 3 complex double per GPU (config 1 scaled)
    dim0 = [32, 32, 32, 64, 4, 3]
    dim1 = [3, 4, 64, 32, 32, 32]
    p0 = np.array([[[0] * 6, dim0]] * 1, dtype=np.int32)
    p1 = np.array([[[0] * 6, dim1]] * 1, dtype=np.int32)
    n = int(np.prod(dim0))
    x, y = field(16, n), torch.zeros(n, device=dev, dtype=torch.complex128)
    record("permute_xyztsc_cstzyx_c128", 2 * n * 16,
           lambda: sb.copy(1, p0, 1, "xyztsc", [0] * 6, dim0, dim0, [x], None, gpu, p1, 1, "cstzyx",
                           [0] * 6, dim1, [y], None, gpu, sb.FastToSlow, sb.Copy))
    if world == 1:
        # (a') the same permutation restricted to the even sites by MaskType masks on both tensors (§8f row 1);
        # bytes as the reference counts them for masked copies: mask size x (sizeof T + sizeof Q), tensor.h:1087
        idx = torch.arange(n, device=dev)
        par = (idx % 32 + (idx // 32) % 32 + (idx // 1024) % 32 + (idx // 32768) % 64) % 2
        m0 = (par == 0).to(torch.float32)
        m1 = m0.view(3, 4, 64, 32, 32, 32).permute(5, 4, 3, 2, 1, 0).contiguous().view(-1)
        del idx, par
        record("masked_even_sites_permute_xyztsc_cstzyx_c128", 2 * n * 16,
               lambda: sb.copy(1, p0, 1, "xyztsc", [0] * 6, dim0, dim0, [x], [m0], gpu, p1, 1, "cstzyx",
                               [0] * 6, dim1, [y], [m1], gpu, sb.FastToSlow, sb.Copy), steps=10)
        del m0, m1
    del x, y
    # (b) periodic +1 shifts of a 64^3 x 128 x 4 x 3 field distributed on z,t (config 5): per GPU block
    pz, pt = grid_for(world)
    for es, tag in ((8, "c64"), (16, "c128")):
        dim = [64, 64, 32 * pz, 32 * pt, 4, 3]  # 64 x 64 x 32 x 32 x (4,3) sites per GPU
        part = sb.basic_partitioning("xyztsc", dim, [1, 1, pz, pt, 1, 1], "zt", world, 1)
        nl = int(np.prod(part[rank, 1]))
        x = field(es, nl)
        y = torch.zeros_like(x)
        for mu, lab in enumerate("xyzt"):
            shift = [0] * 6
            shift[mu] = 1
            record("shift_%s_%s" % (lab, tag), 2 * nl * es,
                   lambda: sb.copy(1, part, 1, "xyztsc", [0] * 6, dim, dim, [x], None, gpu, part, 1,
                                   "xyztsc", shift, dim, [y], None, gpu, sb.FastToSlow, sb.Copy,
                                   comm=comm), steps=10)
        del x, y
    # (c) redistribution t-slabs -> (z,t) blocks of a 32^3 x 64 x (4,3) x n field (config 3), N > 1
    if world > 1:
        dim = [32, 32, 32, 64, 4, 3, 16 * world]
        pa = sb.basic_partitioning("xyztscn", dim, [1, 1, 1, world, 1, 1, 1], "t", world, 1)
        pb = sb.basic_partitioning("xyztscn", dim, [1, 1, pz, pt, 1, 1, 1], "zt", world, 1)
        nl = int(np.prod(pa[rank, 1]))
        x = field(8, nl)
        y = torch.zeros(int(np.prod(pb[rank, 1])), device=dev, dtype=torch.complex64)
        record("redistribute_t_to_zt_c64", 2 * nl * 8,
               lambda: sb.copy(1, pa, 1, "xyztscn", [0] * 7, dim, dim, [x], None, gpu, pb, 1,
                               "xyztscn", [0] * 7, dim, [y], None, gpu, sb.FastToSlow, sb.Copy,
                               comm=comm), steps=10)
        del x, y
    return out


if __name__ == "__main__":
    sys.exit(main())
