"""Aggregate device-to-host bandwidth of N ranks copying concurrently (one rank per GPU), for the
end-to-end scaling analysis: the bench's e2e step is one 440.7 MB D2H per distribution, and on the
8-GPU box eight such copies ran at 95 GB/s in total (round 1).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/tools/d2h_scaling.py

Variants of the host buffer: cudaHostAlloc default / write-combined / portable, and a malloc'ed buffer
registered with cudaHostRegister; copy in one piece and in 32 MB pieces. Prints one JSON line on rank 0."""
import ctypes as C
import glob
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
    glob.glob("/usr/local/cuda/lib64/libcudart.so*")
rt = C.CDLL(cands[0])
N = 440_664_064
dev = torch.empty(N, dtype=torch.uint8, device="cuda").fill_(3)
stream = torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(ptr, piece):
    def once():
        off = 0
        while off < N:
            k = min(piece, N - off)
            rc = rt.cudaMemcpyAsync(C.c_void_p(ptr + off), C.c_void_p(dev.data_ptr() + off), C.c_size_t(k), 2,
                                    C.c_void_p(stream.cuda_stream))
            assert rc == 0, rc
            off += k
        stream.synchronize()
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        once()
    dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"n_gpus": world, "bytes_per_rank": N}
for name, flags in (("default", 0), ("portable", 1), ("write_combined", 4)):
    p = C.c_void_p()
    t0 = time.perf_counter()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(flags))
    alloc_s = time.perf_counter() - t0
    if rc != 0:
        out[name] = {"error": rc}
        continue
    C.memset(p, 0, N)
    for piece_name, piece in (("one_copy", N), ("pieces_32MB", 32 << 20)):
        dt = timed(p.value, piece)
        out[f"{name}/{piece_name}"] = {"ms": dt * 1e3, "aggregate_GBps": world * N / dt / 1e9, "alloc_s": round(alloc_s, 3)}
    rt.cudaFreeHost(p)
buf = (C.c_char * N)()
C.memset(buf, 0, N)
t0 = time.perf_counter()
rc = rt.cudaHostRegister(C.c_void_p(C.addressof(buf)), C.c_size_t(N), C.c_uint(0))
reg_s = time.perf_counter() - t0
if rc == 0:
    dt = timed(C.addressof(buf), N)
    out["registered_malloc/one_copy"] = {"ms": dt * 1e3, "aggregate_GBps": world * N / dt / 1e9, "register_s": round(reg_s, 3)}
    rt.cudaHostUnregister(C.c_void_p(C.addressof(buf)))
try:
    out["numa_nodes"] = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
    out["cpus"] = len(os.sched_getaffinity(0))
except Exception:
    pass
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
