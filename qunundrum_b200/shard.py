"""Sharding of a slice list over worker ranks (one GPU per rank).

Slices are independent units: the reference farms them
out with an MPI master-worker loop in enumerator order
(src/main_generate_distribution.cpp:921-1050, 1139-1362). There is no exchange
step, so no data-path collective: each rank integrates its share, and only the
small per-slice summaries (total probability, total error, flags -- the "slice
histogram" of the distribution) are gathered on rank 0, which is what the
server receives first from every `*_slice_send`.

Slice cost is nearly uniform at a fixed dimension, so a static interleaved partition
of the priority-sorted list balances as well as the dynamic farm and keeps the
dispatch order deterministic. The interleave ROTATES from round to round (slice i ->
rank (i + i // world) mod world): the enumerator lists (alpha_d, alpha_r) and
(-alpha_d, alpha_r) next to each other and the fused kernel's same-sign tile variant
(negative alpha_d) is a few percent cheaper, so a plain i mod world with an even world
size would give the even ranks all the expensive halves (measured at 8 GPUs: 0.162 ms
on the even against 0.153 ms on the odd ranks per step).
"""
from __future__ import annotations

import numpy as np


def partition(n: int, world_size: int, rank: int) -> np.ndarray:
    """Indices (into the enumerator-ordered slice list) owned by `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    i = np.arange(n, dtype=np.int64)
    return i[(i + i // world_size) % world_size == rank]


def enumerate_2d(m: int, t_low: int = 30, t_high: int = 10):
    """The (min_log_alpha_d, min_log_alpha_r) coordinates a generator client
    actually integrates for parameters with t = 30: |alpha| from m - 30 to
    m + 10 (the m + 10 skip rule, src/main_generate_distribution.cpp:1196-1218),
    both signs of alpha_d, alpha_r > 0 (mirrored enumeration,
    src/distribution_enumerator.cpp:56-65), sorted by distance from (m, m) as
    the enumerator's priority sort does (ties broken deterministically here)."""
    coords = []
    for a in range(m - t_low, m + t_high + 1):
        for b in range(m - t_low, m + t_high + 1):
            for sd in (1, -1):
                coords.append((sd * a, b))
    coords.sort(key=lambda c: ((abs(c[0]) - m) ** 2 + (c[1] - m) ** 2, abs(c[0]), c[1], -c[0]))
    return coords


def gather_summaries(local_idx: np.ndarray, local_summary: np.ndarray, n: int,
                     stride: int = 8):
    """All ranks contribute their slices' summaries; rank 0 returns the (n, stride)
    table in list order, the others None. Uses torch.distributed (NCCL on GPUs,
    gloo in the CPU tests); a no-op when not initialised."""
    import torch
    import torch.distributed as dist

    local_summary = np.asarray(local_summary, dtype=np.float64).reshape(-1, stride)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = np.zeros((n, stride))
        out[local_idx] = local_summary
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    per = (n + world - 1) // world
    buf = torch.zeros(per, stride + 1, dtype=torch.float64, device=dev)
    k = len(local_idx)
    buf[:k, 0] = torch.as_tensor(local_idx + 1, dtype=torch.float64, device=dev)  # 0 = padding
    buf[:k, 1:] = torch.as_tensor(local_summary, dtype=torch.float64, device=dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    out = np.zeros((n, stride))
    for b in bufs:
        b = b.cpu().numpy()
        sel = b[:, 0] > 0
        out[(b[sel, 0] - 1).astype(np.int64)] = b[sel, 1:]
    return out
