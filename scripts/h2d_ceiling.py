#!/usr/bin/env python
"""Aggregate host->device copy ceiling of the box: every rank copies a pinned buffer to its GPU in a loop, all ranks at
once.  Plain cudaMemcpyAsync, no kernels.  Default pinned memory and write-combined pinned memory (cudaHostAllocWriteCombined).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/h2d_ceiling.py
"""
import ctypes, json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    host = dist.new_group(backend="gloo")
rt = ctypes.CDLL("libcudart.so.12")
NB = 64 << 20
res = {}
for name, flags in (("pinned", 0), ("pinned_write_combined", 4)):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 1, NB)
    d = torch.empty(NB, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    copy = lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(NB), ctypes.c_int(1), ctypes.c_void_p(st))
    for _ in range(3):
        copy()
    torch.cuda.synchronize()
    for chunk, reps in ((NB, 40), (6 << 20, 200)):
        copyc = lambda: rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), p, ctypes.c_size_t(chunk), ctypes.c_int(1), ctypes.c_void_p(st))
        if world > 1:
            dist.barrier(group=host)
        t0 = time.perf_counter()
        for _ in range(reps):
            copyc()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[f"{name}_{chunk >> 20}MB_GBps_per_gpu"] = chunk * reps / float(dt) / 1e9
        res[f"{name}_{chunk >> 20}MB_GBps_aggregate"] = world * chunk * reps / float(dt) / 1e9
    rt.cudaFreeHost(p)
if rank == 0:
    res["n_gpus"] = world
    res["cpus"] = len(os.sched_getaffinity(0))
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
