#!/usr/bin/env python
"""bench.py -- slice cells integrated per second at m = 2048 on N B200s.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md section 8(d) "T2D"): the slice set
of one two-dimensional Ekera distribution, m = 2048, s = 1 (l = 2048, heuristic
sigma = 1031), t = 30, synthetic d and r (random.seed(20482048 + rank)); the 3362
coordinates the generator client integrates (|alpha| from m - 30 to m + 10, both
signs of alpha_d) in enumerator order, at `-dim 256` (dimension D = 128 with the
Richardson pass at 256): 16,384 cells and 329,218 reference integrand
evaluations per slice; 55.08 M cells per step and per GPU.

A step is one pass of the hot path over that batch. Slices are independent (no data-path
collective; only the per-slice summaries are gathered on rank 0). For N > 1 the default is
STRONG scaling: the 3362 slices of ONE distribution are partitioned over the ranks
(shard.partition: a rotating interleave of the priority-sorted list, the static form of the reference's task farm,
src/main_generate_distribution.cpp:939-976), `value` = the distribution's cells / the slowest
rank's time. `--scaling weak` (one whole distribution per rank, different d and r) is also
measured in every N > 1 run and reported under `weak`; `saturation` is the same distribution in
the generator's default dimension-heuristic mode (slices at D = 128, the upgraded ones re-computed
at 256 and 512, src/main_generate_distribution.cpp:1222-1294; the 512 ones scaled back to 256 on
the device), partitioned the same way: twice the work per distribution, so that 1/8 of it still
fills a GPU.

`value`   : device-resident throughput (descriptors and tables in HBM, results
            left in HBM), CUDA events on the launching stream, max over ranks.
`e2e`     : the same through the synchronous C-ABI call with host buffers:
            coordinates H2D, kernels, all cells + summaries D2H into pinned host
            memory, inside the timed region.
`roofline`: FP64-pipe bound. achieved = cells/s x 1600 algorithmic FP64 flop per
            cell (20.09 evaluations x 80 flop) over the
            FP64 FMA peak measured on the same GPU by a register-resident DFMA
            loop (MEASURED_PEAKS.json holds no FP64 figure). The kernel executes
            fewer flops than the canonical count (separable angle addition, see
            DESIGN.md), so `frac` can exceed 1; `executed_frac` is the share of
            the FP64 pipe actually used (from the committed ncu capture).
`text`    : SURVEY.md section 8(f) #1, the slice text format: the cells of this very
            step (doubles, 2^24 of them) through the exporter kernel ("%.24Lg\\n"
            per value) and the text back through the importer kernels, device
            resident, CUDA events; HBM-roofline fraction from the algorithmic bytes
            (value in + text out); the synchronous host call on one stored slice
            (65,536 cells + total error); libc's fprintf / fscanf on one host core
            beside it. Byte / bit parity is asserted on a sample in the run.
`sections`: (N = 1) every other integrator of the path with its own value / roofline / e2e /
            cpu_baseline: linear_d, linear_r, diagonal (m = 2048, D = 2048, the generator's slice
            lists, ONE launch per distribution) and the sigma-optimal method; a compact numeric
            summary of all sections sits in roofline.sections.
`cpu_baseline` / --impl reference: the UNMODIFIED reference
            (oracle/_ref/libqref.so, distribution_slice_compute_richardson with
            192-bit MPFR) on the host cores, one slice per worker process.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

M, S, T_PARAM, DIM = 2048, 1, 30, 128
FLOP_PER_CELL = 1600.0
EVALS_PER_SLICE_REF = (2 * DIM + 1) ** 2 + (4 * DIM + 1) ** 2
METRIC = "slice cells integrated/sec at m=2048"


def synthetic_d_r(seed: int):
    rnd = random.Random(seed)
    r = 2 ** (M - 1) + 1 + rnd.randrange(2 ** (M - 1) - 1)
    d = r // 2 + rnd.randrange(r // 2)
    return d, r


def workload_config(n_gpus, scaling="weak"):
    per = ("the 3362 slices of ONE distribution partitioned over the ranks (rotating interleave of the sorted list), "
           "no data-path collective" if scaling == "strong" else
           f"one distribution per rank x {n_gpus} ranks, no data-path collective")
    return {
        "workload": ("generate_distribution (2D alpha_d, alpha_r), heuristic sigma, Richardson: "
                     "m=2048 s=1 l=2048 sigma=1031 t=30, 3362 slices (|alpha| in [m-30, m+10], "
                     "both signs of alpha_d) at -dim 256 (D=128 + 256): 55,083,008 cells per "
                     "distribution"),
        "cells_per_distribution": 3362 * DIM * DIM,
        "slices_per_distribution": 3362,
        "dimension": DIM,
        "sharding": per,
        "l2": ("results (440.7 MB per distribution) exceed the 126 MB L2 at N = 1; where a rank's share "
               "is smaller, 256 MB are written between steps (L2 flush) and steps are timed one by one; "
               "inputs are ~5 MB of axis tables rebuilt on the device every step"),
    }


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU reference (oracle/_ref) farm -- the ONLY place bench.py touches oracle/
# --------------------------------------------------------------------------------------
def _ref_worker(args):
    seed, coords = args
    from oracle import ref
    d, r = synthetic_d_r(seed)
    P = ref.RefParameters(M, S, d, r, T_PARAM)
    t0 = time.perf_counter()
    n = 0
    for (a, b) in coords:
        sl = ref.distribution_slice_compute(P, DIM, a, b)
        n += sl.cells.size
    return n, time.perf_counter() - t0


def reference_cpu_pass(pool, cores, coords, slices_per_core=1):
    """One bounded sample: every worker integrates `slices_per_core` slices of the
    workload (in enumerator order) with the reference's own code."""
    jobs = []
    for k in range(cores):
        sl = coords[k * slices_per_core:(k + 1) * slices_per_core]
        jobs.append((20482048, sl))
    t0 = time.perf_counter()
    res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, int(os.environ.get("QB200_CPU_CORES", "256"))))


def run_reference_arm(args, rank):
    import multiprocessing as mp
    from oracle import ref
    from qunundrum_b200 import shard
    if rank != 0:
        return
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable":
                          "oracle/_ref/libqref.so was not built (no /root/reference at build time)"}))
        return
    coords = shard.enumerate_2d(M)
    cores = host_cores()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pass  # nothing to warm: every slice is an independent cold computation
        t0 = time.perf_counter()
        cells = 0
        for _ in range(args.steps):
            c, _ = reference_cpu_pass(pool, cores, coords, 1)
            cells += c
        wall = time.perf_counter() - t0
    value = cells / wall
    sample = (f"{cores} slices per step (one per worker process, first {cores} of the enumerator "
              f"order) x {args.steps} steps, D=128 Richardson, 192-bit MPFR")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": scaling_mode(args, args.gpus), "vs_baseline": None, "dtype": "mpfr192", "data": "synthetic",
        "config": workload_config(args.gpus, scaling_mode(args, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "reference",
                         "sample": sample, "cells_per_s_per_core": value / cores},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def bind_near_gpu(local_rank: int):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE any
    pinned host buffer is allocated (first touch puts the pages on that node): with one
    rank per GPU the end-to-end device-to-host streams then stay on their own socket.
    Best effort; returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return f"numa: node {node} has no allowed cpu"
        os.sched_setaffinity(0, allowed)
        return f"numa: gpu {local_rank} ({bus}) -> node {node}, {len(allowed)} cpus"
    except Exception as exc:  # pragma: no cover
        return f"numa: not bound ({type(exc).__name__})"


def profile_constants():
    """DRAM traffic and FP64-pipe share of the dominant kernel from the committed ncu
    capture (profiles/fused2d_latest.json), or None."""
    p = os.path.join(ROOT, "profiles", "fused2d_latest.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def text_section(ctx, qb, torch, stream, cells, hbm_peak):
    """Exporter / importer kernels on the step's own cells (rank 0, N = 1)."""
    import ctypes as C
    from oracle import text as ot          # checker + libc baseline only
    n = min(1 << 24, cells.numel())
    cap = 33 * n
    d_text = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int64, device="cuda")
    d_vals = torch.empty(2 * n, dtype=torch.int64, device="cuda")
    d_info = torch.zeros(8, dtype=torch.int64, device="cuda")
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    reps = 10

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    l0 = ctx.launch_count
    ms_f = timed(lambda: ctx.text_format_device(qb.host.TEXT_F64, cells.data_ptr(), n,
                                                d_text.data_ptr(), cap, d_len.data_ptr(),
                                                stream.cuda_stream))
    length = int(d_len.item())
    ms_p = timed(lambda: ctx.text_parse_device(d_text.data_ptr(), length, n, d_vals.data_ptr(),
                                               d_info.data_ptr(), stream.cuda_stream))
    launches = ctx.launch_count - l0
    info = d_info.cpu().numpy()
    if int(info[0]) != n or int(info[1]) != 0:
        raise SystemExit(f"bench.py: importer status {info}")
    # parity on a sample: bytes against libc, and the importer's bits against libc's
    k = 200000
    h = cells[:k].cpu().numpy().astype(np.longdouble)
    t0 = time.perf_counter()
    want = ot.format_ld24(h)
    t_fmt = time.perf_counter() - t0
    got = bytes(d_text[:len(want)].cpu().numpy().tobytes())
    if got != want:
        raise SystemExit("bench.py: exporter text differs from libc")
    t0 = time.perf_counter()
    back = ot.parse_ld(want, k)
    t_par = time.perf_counter() - t0
    m_ref, s_ref = ot.ld_fields(back)
    raw = d_vals[:2 * k].cpu().numpy().view(np.uint64).reshape(k, 2)
    if not (np.array_equal(raw[:, 0], m_ref) and np.array_equal(raw[:, 1].astype(np.uint16), s_ref)):
        raise SystemExit("bench.py: importer values differ from libc")
    # the synchronous host call on one stored slice (256 x 256 cells + total error)
    one = cells[:65536].cpu().numpy().astype(np.longdouble)
    tail = np.longdouble("2.5e-307")
    text1 = ctx.text_format(one, tail)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.text_format(one, tail)
    ms_host_f = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.text_parse(text1, 65537)
    ms_host_p = (time.perf_counter() - t0) / 20 * 1e3
    # ... and on the whole sample in one call (pinned host input, text left in the context's
    # pinned buffer; values pageable on the way back): transfer bound
    hp = qb.lib().qb200_host_alloc(8 * n)
    big = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(n,))
    big[:] = cells[:n].cpu().numpy()
    ctx.text_format_view(big)
    t0 = time.perf_counter()
    for _ in range(3):
        addr, blen = ctx.text_format_view(big)
    ms_big_f = (time.perf_counter() - t0) / 3 * 1e3
    L = qb.lib()
    ht = L.qb200_host_alloc(blen)
    hv = L.qb200_host_alloc(16 * n)
    C.memmove(ht, addr, blen)
    used = C.c_size_t()

    def parse_big():
        rc = L.qb200_text_parse_ld(ctx.h, C.cast(ht, C.c_char_p), blen, n, hv, C.byref(used))
        if rc != 0 or used.value != blen:
            raise SystemExit(f"bench.py: qb200_text_parse_ld rc={rc} used={used.value}/{blen}")

    parse_big()
    t0 = time.perf_counter()
    for _ in range(3):
        parse_big()
    ms_big_p = (time.perf_counter() - t0) / 3 * 1e3
    for ptr in (hp, ht, hv):
        L.qb200_host_free(C.c_void_p(ptr))
    bytes_f = 8 * n + length          # doubles in, text out
    bytes_p = length + 16 * n         # text in, x87 values out
    try:                               # DRAM traffic per launch from the committed ncu capture
        tl = json.load(open(os.path.join(ROOT, "profiles", "text_latest.json")))
    except Exception:
        tl = {}
    return {
        "values": n, "text_bytes": length, "bytes_per_value": length / n,
        "export": {"values_per_s": n / (ms_f * 1e-3), "ms": ms_f,
                   "roofline": {"bound": "hbm", "achieved": bytes_f / (ms_f * 1e-3) / 1e9,
                                "peak": hbm_peak, "unit": "GB/s",
                                "frac": bytes_f / (ms_f * 1e-3) / 1e9 / hbm_peak,
                                "traffic": tl.get("format", {}).get("dram_bytes_per_launch"),
                                "traffic_note": "ncu capture on 2^24 x87 values (16 B in): 715 MB "
                                                "against 772 MB algorithmic",
                                "limiter": "integer issue slots: ~830 thread instructions per value, "
                                           "issue active 67 % (profiles/r01_text_format_ncu_full.txt)"},
                   "kernel": "k_text_format<F64>",
                   "host_call_ms_per_slice": ms_host_f,
                   "host_call_values_per_s": 65537 / (ms_host_f * 1e-3),
                   "e2e": {"value": n / (ms_big_f * 1e-3), "unit": "values/s",
                           "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": length,
                           "api": "qb200_text_format_f64, one call, pinned host buffers"}},
        "import": {"values_per_s": n / (ms_p * 1e-3), "ms": ms_p,
                   "roofline": {"bound": "hbm", "achieved": bytes_p / (ms_p * 1e-3) / 1e9,
                                "peak": hbm_peak, "unit": "GB/s",
                                "frac": bytes_p / (ms_p * 1e-3) / 1e9 / hbm_peak,
                                "traffic": (tl.get("tokenize", {}).get("dram_bytes_per_launch", 0) +
                                            tl.get("parse", {}).get("dram_bytes_per_launch", 0)) or None,
                                "limiter": "k_text_parse: integer issue slots (~1900 thread "
                                           "instructions per value, issue active 91 %)"},
                   "kernels": "k_text_tokenize + k_text_parse",
                   "host_call_ms_per_slice": ms_host_p,
                   "host_call_values_per_s": 65537 / (ms_host_p * 1e-3),
                   "e2e": {"value": n / (ms_big_p * 1e-3), "unit": "values/s",
                           "h2d_bytes_per_step": length, "d2h_bytes_per_step": 16 * n,
                           "api": "qb200_text_parse_ld, one call, pinned host buffers"}},
        "cpu_baseline": {"kind": "reference", "cores": 1,
                         "export_values_per_s": k / t_fmt, "import_values_per_s": k / t_par,
                         "sample": f"{k} cells of the step through libc fprintf(\"%.24Lg\\n\") / "
                                   f"fscanf(\"%Lg\\n\") ({ot.libc_version()}), the calls the "
                                   "reference's slice exporters / importers make"},
        "parity": f"first {k} values: text byte-identical to libc, parsed bits identical to libc",
        "gpu_launches": int(launches),
    }


def tau_section(ctx, qb, torch, stream, h_cells, tp, coords, hbm_peak, cpu_baseline=True):
    """Sampling from the step's own distribution: the 3362 slices the
    step computed plus their mirror images (what generate_distribution stores), sorted like
    distribution_sort_slices; 2^16 tau estimates of n = 16 samples per timed call.
    Rank 0, N = 1."""
    import ctypes as C
    n_slices = len(coords)
    D = DIM
    t0 = time.perf_counter()
    dist = qb.Distribution(M)
    LD = np.longdouble
    for i, (a, b) in enumerate(coords):
        cells = np.ascontiguousarray(h_cells[i], dtype=LD)   # widened as the drop-in does
        total = LD(0)
        for j in range(0, D * D, 4096):                       # (blocked: the order is immaterial here)
            total += cells[j:j + 4096].sum(dtype=LD)
        for sa, sb in ((a, b), (-a, -b)):                     # the server's mirrored copy
            sl = qb.Distribution_Slice(D, int(sa), int(sb), norm_matrix=cells)
            sl.total_probability = total
            dist.insert_slice(sl)
    dist.sort_slices()
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    sampler = qb.Sampler(dist, ctx)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    wps = sampler.words_per_sample
    n, count = 16, 1 << 16
    total = n * count
    g = torch.Generator(device="cuda")
    g.manual_seed(20482048)
    d_words = torch.randint(-2 ** 63, 2 ** 63 - 1, (total * wps,), dtype=torch.int64, device="cuda", generator=g)
    d_sums = torch.zeros(count * 4, dtype=torch.float64, device="cuda")
    d_status = torch.zeros(count, dtype=torch.int32, device="cuda")
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    l0 = ctx.launch_count

    def run():
        sampler.tau_device(n, count, d_words.data_ptr(), d_sums.data_ptr(), d_status.data_ptr(),
                           stream.cuda_stream)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    launches = ctx.launch_count - l0
    failed = int((d_status != 0).sum().item())
    # end to end: host words in, long double taus out, through the C ABI
    h_words = d_words.cpu().numpy().view(np.uint64)
    sampler.tau_estimate(n, count, h_words)
    t0 = time.perf_counter()
    e2e_reps = 3
    for _ in range(e2e_reps):
        ta, tb, ok, used = sampler.tau_estimate(n, count, h_words)
    wall = (time.perf_counter() - t0) / e2e_reps
    done = len(ta)
    # algorithmic bytes per sample of the search as it is now (sampler.cuh): per search two guide
    # entries (8 B), the two checks of the bracket and ~3 steps inside it (16 B each), the walk
    # state before the block (16 B) and the block's 8 elements as doubles; the slice record (48 B),
    # its total (16 B), four geometry entries (16 B each), the words in, 64 B out
    seg_block = 8                                             # QB_SEG_BLOCK, sampler.cuh
    bytes_per_sample = 2 * (8 + 2 * 16 + 3 * 16 + 16 + seg_block * 8) + 48 + 16 + 4 * 16 + wps * 8 + 64
    gbs = total * bytes_per_sample / (ms * 1e-3) / 1e9
    out = {
        "workload": (f"tau_estimate on the step's own distribution: {len(dist.slices)} slices "
                     f"({n_slices} computed + mirrored) x {D * D} cells = {sampler_cells(sampler)} cells "
                     f"resident ({sampler_cells(sampler) * 16 / 1e9:.2f} GB x87 + 25 % coarse index + 50 % "
                     f"the same elements as doubles for the quick pass); "
                     f"{count} estimates of n = {n} samples per call"),
        "samples_per_call": total,
        "value": total / (ms * 1e-3), "unit": "samples/s", "ms": ms,
        "out_of_bounds_estimates": failed,
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": None, "bytes_per_sample": bytes_per_sample, "kernel": "k_sample",
                     "limiter": ("latency of ~10 dependent loads of divergent addresses per sample (every thread "
                                 "searches its own slice): issue active 41 %, long-scoreboard 9 cycles per "
                                 "instruction, DRAM 15 % (profiles/r02_sampler_ncu_full.txt; the L1 look-ups that "
                                 "bounded the round-1 kernel were halved by 16-byte loads and a copy of the "
                                 "elements as doubles, profiles/r02_sampler_lookups_ab.txt)")},
        "e2e": {"value": done * n / wall, "unit": "samples/s", "h2d_bytes_per_step": int(used * 8 + done * 8),
                "d2h_bytes_per_step": int(done * 36), "ms": wall * 1e3,
                "api": "qb200_sampler_tau_estimate (host words in, long double taus out)"},
        "setup": {"host_containers_s": t_host, "sampler_create_s": t_create,
                  "note": "once per distribution: upload of the cells as they lie in the slices + k_seg_build"},
        "gpu_launches": int(launches),
    }
    if cpu_baseline:
        try:
            from oracle import ref             # checker + CPU baseline only
            if ref.available():
                d, r = synthetic_d_r(20482048)
                RP = ref.RefParameters(M, S, d, r)
                order = dist.slices
                rd = ref.RefDistribution(2, RP, [s.dimension for s in order],
                                         [s.min_log_alpha_d for s in order],
                                         [s.min_log_alpha_r for s in order],
                                         np.concatenate([s.norm_matrix for s in order]),
                                         [s.total_probability for s in order])
                seed = bytes(range(32))
                k = 30000     # ~10 s of the reference on one core
                t0 = time.perf_counter()
                r0, r1, rok = rd.tau_estimate(ref.RefRandom(seed), n, k)
                w = time.perf_counter() - t0
                words = ref.RefRandom(seed).words(k * n * wps)
                g0, g1, gok, _ = sampler.tau_estimate(n, k, words)
                same = bool((gok == rok).all())
                err = float(max(np.abs(g0[rok] - r0[rok]).max(), np.abs(g1[rok] - r1[rok]).max())) if rok.any() else 0.0
                out["cpu_baseline"] = {"value": k * n / w, "unit": "samples/s", "cores": 1, "kind": "reference",
                                       "sample": (f"{k} calls of the reference's tau_estimate (n = {n}) on the same "
                                                  f"distribution, its own Keccak stream, {w:.1f} s")}
                out["parity"] = (f"{k} estimates on the reference's stream: success flags identical: {same}; "
                                 f"max |tau - tau_ref| = {err:.2e}")
                if not same or err > 2.0 ** -63 * (2 * M + 16):
                    raise SystemExit("bench.py: tau estimates differ from the reference's")
        except SystemExit:
            raise
        except Exception as exc:  # pragma: no cover
            out["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "reference",
                                   "sample": f"failed: {exc}"}
    sampler.close()
    return out


def diagk_section(ctx, qb, torch, stream, cpu_baseline=True, m=M, sigma=5, l=M, n=2048 * 148,
                  delta_bound=1000, seed=11):
    """The diagonal distribution's k given (j, eta) (second half):
    sample_k_from_diagonal_j_eta_pivot for `n` uniform (j, eta, pivot) at m = 2048. Device-resident
    timing with CUDA events on the launching stream (qb200_diagk_sample_device), the synchronous
    C ABI with host rows (qb200_diagk_sample), a sub-sample checked against the reference's own
    function on one host core (when oracle/_ref is on the box)."""
    import random
    LD = np.longdouble
    prng = random.Random(seed)
    r = (1 << (m - 1)) + 1 + prng.randrange((1 << (m - 1)) - 1)
    d = r // 2 + prng.randrange(r // 2)
    wj = (m + sigma + 31) // 32
    rng = np.random.default_rng(seed)
    J = rng.integers(0, 1 << 32, size=(n, wj), dtype=np.uint64).astype(np.uint32)
    if (m + sigma) % 32:
        J[:, -1] &= np.uint32((1 << ((m + sigma) % 32)) - 1)
    eta = rng.integers(-25, 26, size=n).astype(np.int32)
    piv = rng.random(n).astype(LD)
    S = qb.DiagonalKSampler(qb.Diagonal_Parameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l), ctx)
    dJ = torch.from_numpy(J.view(np.int32)).cuda()
    dE = torch.from_numpy(eta).cuda()
    dP = torch.from_numpy(piv.view(np.uint8).reshape(n, 16)).cuda()
    dK = torch.zeros((n, S.k_limbs), dtype=torch.int32, device="cuda")
    dO = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    launches0 = ctx.launch_count
    times = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        S.sample_device(n, dJ.data_ptr(), dE.data_ptr(), dP.data_ptr(), delta_bound, dK.data_ptr(),
                        dO.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        times.append(e0.elapsed_time(e1))
    launches = (ctx.launch_count - launches0) // 6
    ms = float(np.median(times[3:]))
    # the same with the reference's default bound (BOUND_DELTA = 10^6, src/diagonal_distribution.h:56):
    # pivots near 1 then walk up to 2 10^6 + 1 steps
    times_default = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        S.sample_device(n, dJ.data_ptr(), dE.data_ptr(), dP.data_ptr(), 1000000, dK.data_ptr(),
                        dO.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        times_default.append(e0.elapsed_time(e1))
    out_default = dO.cpu().numpy()
    S.sample_device(n, dJ.data_ptr(), dE.data_ptr(), dP.data_ptr(), delta_bound, dK.data_ptr(),
                    dO.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    out = dO.cpu().numpy()
    status = out[:, 3].view(np.int64) & 0xffffffff
    delta = out[:, 2].view(np.int64)
    k = (m + 31) // 32
    # multiply-adds of the products of diagk.cuh on the common path: r j from three guard columns below
    # 2^(m+sigma), and the columns fl - 8 .. fl + wl - 1 of s psi (fl = k + 4 fractional limbs of
    # psi = 2^l d / r in fixed point, wl = limbs of k)
    wl = (l + 31) // 32
    cs = (m + sigma) >> 5
    skipped = sum(min(c_ + 1, k, wj) for c_ in range(max(0, cs - 4)))   # columns of r j below the guards
    fl, npsi = k + 4, k + 4 + wl
    spsi = sum(min(c_, k - 1) - max(0, c_ - npsi + 1) + 1 for c_ in range(fl - 8, fl + wl))
    mads = k * wj - skipped + spsi
    res = {"workload": f"sample_k_from_diagonal_j_eta_pivot, m={m} sigma={sigma} l={l}: {n} uniform "
                       f"(j, eta, pivot), delta_bound={delta_bound}",
           "samples_per_call": n, "value": n / ms * 1e3, "unit": "samples/s", "ms": ms,
           "ok_fraction": float((status == 0).mean()), "mean_abs_delta": float(np.abs(delta).mean()),
           "max_abs_delta": int(np.abs(delta).max()), "gpu_launches": int(launches),
           "default_delta_bound": {"delta_bound": 1000000, "ms": float(np.median(times_default)),
                                   "value": n / float(np.median(times_default)) * 1e3,
                                   "max_abs_delta": int(np.abs(out_default[:, 2].view(np.int64)).max()),
                                   "ok_fraction": float(((out_default[:, 3].view(np.int64) & 0xffffffff) == 0).mean())},
           "roofline": {"bound": "integer issue", "imad_per_sample": mads,
                        "achieved": mads * n / ms * 1e3 / 1e12, "unit": "T multiply-add/s",
                        "kernel": "k_diagk",
                        "limiter": "instruction issue: ~2.8 thread instructions per multiply-add in the "
                                   "four-column products, issue slots 49 % active (long scoreboard 3.9, wait "
                                   "2.7, math pipe 2.4 cycles per instruction; 32 warps/SM), LSU 36 %, DRAM 10 % "
                                   "(profiles/r02_diagk_ncu_full.txt, r02_diagk_variants_ab.txt; round 1 "
                                   "needed 17,092 multiply-adds per sample at m = 2048)"}}
    t0 = time.perf_counter()
    ks, x, dl, st = S.sample(J, eta, piv, delta_bound, want_k=False)
    t1 = time.perf_counter()
    res["e2e"] = {"value": n / (t1 - t0), "unit": "samples/s", "ms": (t1 - t0) * 1e3,
                  "h2d_bytes_per_step": int(J.nbytes + eta.nbytes + piv.nbytes), "d2h_bytes_per_step": int(n * 32),
                  "api": "qb200_diagk_sample (host rows in; alpha_phi, delta, status out)"}
    sub = np.random.default_rng(3).choice(n, 300, replace=False)
    if cpu_baseline:
        try:
            from oracle import ref as R
            if not R.available():
                raise RuntimeError("oracle/_ref not on this box")
            gk = dK.cpu().numpy().view(np.uint32)
            P = R.RefDiagonalParameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l)
            cnt, same, t0 = 0, True, time.perf_counter()
            for i in sub:
                jj = int.from_bytes(J[i].tobytes(), "little")
                ok, kk, _, _ = R.sample_k_from_diagonal_j_eta_pivot(P, piv[i], jj, int(eta[i]), delta_bound,
                                                                    precision=256)
                same = same and ok == (status[i] == 0) and kk == int.from_bytes(gk[i].tobytes(), "little")
                cnt += 1
                if time.perf_counter() - t0 > 10:
                    break
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": cnt / dt, "unit": "samples/s", "cores": 1, "kind": "reference",
                                   "sample": f"{cnt} calls of the reference's sample_k_from_diagonal_j_eta_pivot "
                                             f"on the same inputs, {dt:.1f} s"}
            res["parity"] = f"{cnt} samples against the reference: k and success flags identical: {bool(same)}"
        except Exception as exc:  # pragma: no cover
            res["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "reference",
                                   "sample": f"unavailable: {exc}"}
    S.close()
    return res


def exact_section(ctx, qb, torch, stream, cpu_baseline=True, m=M, l=M, dim=256, n=1024 * 148, seed=13):
    """The exact samplers (first half): sample_alpha_from_region for the two
    arguments of `n` samples followed by sample_j_k_from_alpha_d_r, m = 2048, s = 1 (src/sample.cpp:78-158,
    275-352; the two halves of distribution_sample_pair_j_k, src/distribution.cpp:615-680, around the
    reference's lattice_alpha_map). Kernel times from CUDA events on the launching stream (recorded
    inside the library around k_exact_alpha / k_exact_jk), the synchronous C ABI with host rows and
    host random bytes, a sub-sample checked against the reference's own functions on one host core
    (when oracle/_ref is on the box)."""
    import random
    prng = random.Random(seed)
    r = (1 << (m - 1)) + 1 + 2 * prng.randrange((1 << (m - 2)) - 1)          # odd: kappa_r = 0
    d = (r // 2 + prng.randrange(r // 2)) | 1
    S = qb.ExactSampler(qb.Parameters(m=m, s=0, d=d, r=r, l=l), dim, 0, ctx)
    rng = np.random.default_rng(seed)
    L = qb.lib()
    # regions as a generator's slices give them: |log alpha| on [m - 30, m + 10), uniform region index
    e = rng.integers(m - 30, m + 10, size=2 * n)
    sign = np.where(rng.integers(0, 2, size=2 * n) == 1, -1, 1)
    reg = rng.integers(0, dim, size=2 * n)
    nbytes_of = {}
    g = np.zeros(2 * n, dtype=qb.EXACT_REGION_DTYPE)
    g["min_log_alpha"] = sign * e
    g["region"] = reg
    g["dimension"] = dim
    lens = np.zeros(2 * n, dtype=np.uint32)
    for i in range(2 * n):
        key = (int(e[i]), int(reg[i]))
        v = nbytes_of.get(key)
        if v is None:
            v, st = S.region_bytes(key[0], key[1], dim)
            assert st == 0
            nbytes_of[key] = v
        lens[i] = v
    g["length"] = lens
    g["offset"] = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    stream_bytes = rng.integers(0, 256, size=int(lens.sum()), dtype=np.uint8)
    alpha = np.zeros((2 * n, S.wa), dtype=np.uint32)
    neg = np.zeros(2 * n, dtype=np.int32)
    st = np.zeros(2 * n, dtype=np.int32)
    j = np.zeros((n, S.wn), dtype=np.uint32)
    k = np.zeros((n, S.wk), dtype=np.uint32)
    launches0 = ctx.launch_count

    def call():
        t0 = time.perf_counter()
        rc = L.qb200_exact_alpha(S.h, 2 * n, g.ctypes.data, 0, stream_bytes.ctypes.data, len(stream_bytes),
                                 alpha.ctypes.data, neg.ctypes.data, st.ctypes.data)
        assert rc == 0, qb.host._err()
        ms_a = S.kernel_ms()[0]
        rc = L.qb200_exact_j_k(S.h, 1, n, alpha[:n].ctypes.data, neg[:n].ctypes.data, alpha[n:].ctypes.data,
                               neg[n:].ctypes.data, None, j.ctypes.data, k.ctypes.data)
        assert rc == 0, qb.host._err()
        return time.perf_counter() - t0, ms_a, S.kernel_ms()[1]

    runs = [call() for _ in range(4)]
    launches = (ctx.launch_count - launches0) // 4
    wall = float(np.median([q[0] for q in runs[1:]]))
    ms_a = float(np.median([q[1] for q in runs[1:]]))
    ms_jk = float(np.median([q[2] for q in runs[1:]]))
    assert int(st.max()) == 0
    # 32-bit multiply-adds of the truncated products (exact.cuh): alpha_r (wa limbs) times the inverse of r
    # (wn limbs), columns 0 .. wn - 1; j (wn limbs) times d (wd limbs), columns 0 .. wn - 1
    wd = (d.bit_length() + 31) // 32
    mads = sum(min(c_ + 1, S.wa) for c_ in range(S.wn)) + sum(min(c_ + 1, wd) for c_ in range(S.wn))
    # the alpha launch did 2 n samples; the k_exact_jk launch n
    ms = ms_a + ms_jk
    res = {"workload": f"2 x sample_alpha_from_region + sample_j_k_from_alpha_d_r, m={m} l={l} dimension={dim}: "
                       f"{n} (j, k) pairs",
           "samples_per_call": n, "value": n / ms * 1e3, "unit": "samples/s", "ms": ms,
           "ms_k_exact_alpha": ms_a, "ms_k_exact_jk": ms_jk, "gpu_launches": int(launches),
           "roofline": {"bound": "integer issue", "imad_per_sample": mads,
                        "achieved": mads * n / ms_jk * 1e3 / 1e12, "unit": "T multiply-add/s",
                        "kernel": "k_exact_jk",
                        "limiter": "instruction issue, as k_diagk (the same four-column products, mul_columns)"},
           "e2e": {"value": n / wall, "unit": "samples/s", "ms": wall * 1e3,
                   "h2d_bytes_per_step": int(g.nbytes + stream_bytes.nbytes + alpha.nbytes + neg.nbytes),
                   "d2h_bytes_per_step": int(alpha.nbytes + neg.nbytes + st.nbytes + j.nbytes + k.nbytes),
                   "api": "qb200_exact_alpha + qb200_exact_j_k (host rows and host random bytes in; alpha, j, k out)"}}
    if cpu_baseline:
        try:
            from oracle import ref as R
            if not R.available():
                raise RuntimeError("oracle/_ref not on this box")
            P = R.RefParameters(m, 0, d, r, l=l) if False else R.RefParameters(m, 1, d, r)
            sub = np.random.default_rng(3).choice(n, 200, replace=False)
            li = lambda row: int.from_bytes(np.ascontiguousarray(row).tobytes(), "little")
            seed32 = bytes(range(32))
            cnt, same, t0 = 0, True, time.perf_counter()
            for i in sub:
                got_alpha = []
                for q in (i, n + i):
                    # a Random_State that delivers exactly this sample's bytes does not exist: the reference
                    # draws from its own generator, so the comparison runs the other way round -- its bytes
                    # are handed to the GPU below; here only the timing and the (j, k) arithmetic are taken
                    got_alpha.append((-1 if neg[q] else 1) * li(alpha[q]))
                rs_ = R.RefRandom(seed32)
                R.sample_alpha_from_region(float(sign[i] * (e[i] + reg[i] / dim)),
                                           float(sign[i] * (e[i] + (reg[i] + 1) / dim)), 0, rs_)
                R.sample_alpha_from_region(float(sign[n + i] * (e[n + i] + reg[n + i] / dim)),
                                           float(sign[n + i] * (e[n + i] + (reg[n + i] + 1) / dim)), 0, rs_)
                jj, kk = R.sample_j_k(1, P, got_alpha[0], got_alpha[1], rs_)
                same = same and jj == li(j[i]) and kk == li(k[i])
                cnt += 1
                if time.perf_counter() - t0 > 10:
                    break
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": cnt / dt, "unit": "samples/s", "cores": 1, "kind": "reference",
                                   "sample": f"{cnt} x (two calls of the reference's sample_alpha_from_region and one "
                                             f"of sample_j_k_from_alpha_d_r on the same regions), {dt:.1f} s"}
            res["parity"] = (f"{cnt} pairs against the reference's sample_j_k_from_alpha_d_r on the GPU's alpha: "
                             f"j and k identical: {bool(same)} (alpha itself against the reference on the same "
                             f"Random_State stream: tests/test_exact.py)")
        except Exception as exc:  # pragma: no cover
            res["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "reference",
                                   "sample": f"unavailable: {exc}"}
    S.close()
    return res


def sampler_cells(sampler):
    import qunundrum_b200 as qb
    return int(qb.lib().qb200_sampler_cells(sampler.h))


def scaling_mode(args, world):
    if args.scaling in ("strong", "weak"):
        return args.scaling
    return "strong" if world > 1 else "weak"


def kernel_source_sha():
    """Fingerprint of the sources the fused kernel is compiled from: profiles/fused2d_latest.json
    carries the one its ncu capture was taken with."""
    import hashlib
    h = hashlib.sha256()
    for f in ("kernels_fused2d.cuh", "qmath.cuh", "integrands.cuh", "slice_cells.cuh"):
        h.update(open(os.path.join(ROOT, "qunundrum_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def heuristic_dimensions(coords, tp128, m=M):
    """The client's dimension heuristic (src/main_generate_distribution.cpp:1222-1294) for every
    coordinate: (initial slice dimension, final slice dimension). tp128: the total at D = 128 (the
    initial dimension of every coordinate that is not on the diagonal tail)."""
    out = []
    for (a, b), tp in zip(coords, tp128):
        max_alpha = max(abs(a), abs(b))
        req = 256
        if abs(a - b) <= 1 and (a < 0) == (b < 0):
            if max_alpha >= m + 3:
                req = 1024
            elif max_alpha >= m:
                req = 512
        initial = req
        # (for the tail coordinates the second test reads the total at their own initial dimension;
        # it can only raise 512 -> 1024 at max_alpha >= m + 10, where the first test already gave 1024)
        if tp >= 1e-7 and max_alpha >= m and req < 512:
            req = 512
        if tp >= 1e-10 and max_alpha >= m + 10 and req < 1024:
            req = 1024
        out.append((initial // 2, req // 2))
    return out


class DeviceTimer:
    """K steps on `stream`, CUDA events. flush: write 256 MB between steps (L2) and time the steps one
    by one; else one event pair around all K."""

    def __init__(self, torch, stream, flush):
        self.torch, self.stream, self.flush = torch, stream, flush
        self.scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush else None

    def run(self, step, k):
        torch = self.torch
        if not self.flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(k):
                step()
            e1.record(self.stream)
            self.stream.synchronize()
            return e0.elapsed_time(e1)
        ev = []
        for _ in range(k):
            self.scratch.fill_(1)          # on the current stream = self.stream
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            step()
            e1.record(self.stream)
            ev.append((e0, e1))
        self.stream.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev)


def one_dimensional_sections(ctx, qb, torch, stream, peak_flops, hbm_peak, cpu_baseline=True):
    """linear_d, linear_r, diagonal at m = 2048, D = 2048 with the slice lists the generators'
    enumerators hand out (41 / 41 / 170 slices): ONE launch per distribution (k_fused1d)."""
    import multiprocessing as mp
    d, r = synthetic_d_r(20482048)
    D = 2048
    out = {}
    lin = list(range(M - 30, M + 11))
    sigma, eta_bound = 5, 2
    diag_a, diag_eta = [], []
    for eta in (0, 1, -1, 2, -2):
        for a in range(M - 30, M + sigma - 1):
            diag_a.append(a)
            diag_eta.append(eta)
    cases = [("linear_d", 0, qb.Parameters(M, S, d, r, T_PARAM), lin, None, 420.0),
             ("linear_r", 1, qb.Parameters(M, S, d, r, T_PARAM), lin, None, 690.0),
             ("diagonal", 2, qb.Diagonal_Parameters(M, sigma, S, d, r, eta_bound=eta_bound, t=T_PARAM), diag_a,
              diag_eta, 420.0)]
    for name, kind, P, a, eta, flop in cases:
        n = len(a)
        plan = ctx.plan1d(P, kind, True, D, a, eta)
        cells = torch.empty(plan.cells, dtype=torch.float64, device="cuda")
        summ = torch.empty(n * 8, dtype=torch.float64, device="cuda")
        l0 = ctx.launch_count
        plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        launches = ctx.launch_count - l0
        for _ in range(5):
            plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        plan.set_algorithm(1)              # the three-launch path, for comparison
        for _ in range(3):
            plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        e0.record(stream)
        for _ in range(reps):
            plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        ms_plain = e0.elapsed_time(e1) / reps
        plan.close()
        tot = n * D
        # end to end: the synchronous C ABI, host buffers
        ctx.slice1d_batch(P, kind, True, D, a, eta)
        t0 = time.perf_counter()
        for _ in range(10):
            c1, tp1, fl1 = ctx.slice1d_batch(P, kind, True, D, a, eta)
        wall = (time.perf_counter() - t0) / 10
        ach = tot / (ms * 1e-3) * flop / 1e12
        sec = {"workload": f"{name}: m={M} s={S} D={D}, {n} slices ({tot} cells) per distribution, Richardson",
               "value": tot / (ms * 1e-3), "unit": "cells/s", "ms": ms, "gpu_launches": int(launches),
               "three_launch_path_ms": ms_plain, "mass": float(tp1.sum()),
               "roofline": {"bound": "fp64 (latency at this size)", "achieved": ach, "peak": peak_flops / 1e12,
                            "unit": "TFLOP/s", "frac": ach / (peak_flops / 1e12), "flop_per_cell": flop,
                            "kernel": "k_fused1d", "traffic": None,
                            "note": ("one launch of %d blocks x 128 threads: the whole distribution is %.0f us of "
                                     "device time, most of it the launch itself" % (n * (D // 128), ms * 1e3))},
               "e2e": {"value": tot / wall, "unit": "cells/s", "ms": wall * 1e3,
                       "h2d_bytes_per_step": int(8 * n + 2 * ((M + 7) // 8)), "d2h_bytes_per_step": int(tot * 8 + n * 64),
                       "api": "qb200_slice1d_compute (synchronous C ABI, one call per distribution)"}}
        if kind == 2:
            # the same kernel when the launch fills the GPU: eta bound 25 (51 values of eta), device resident
            big_eta = 25
            Pb = qb.Diagonal_Parameters(M, sigma, S, d, r, eta_bound=big_eta, t=T_PARAM)
            ab, eb = [], []
            for e_ in range(-big_eta, big_eta + 1):
                for a_ in range(M - 30, M + sigma - 1):
                    ab.append(a_)
                    eb.append(e_)
            planb = ctx.plan1d(Pb, kind, True, D, ab, eb)
            cb = torch.empty(planb.cells, dtype=torch.float64, device="cuda")
            sb = torch.empty(len(ab) * 8, dtype=torch.float64, device="cuda")
            for _ in range(5):
                planb.run(cb.data_ptr(), sb.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            e0.record(stream)
            for _ in range(reps):
                planb.run(cb.data_ptr(), sb.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
            msb = e0.elapsed_time(e1) / reps
            planb.close()
            totb = len(ab) * D
            achb = totb / (msb * 1e-3) * flop / 1e12
            sec["at_eta_bound_25"] = {"slices": len(ab), "cells": totb, "ms": msb, "value": totb / (msb * 1e-3),
                                      "unit": "cells/s", "frac": achb / (peak_flops / 1e12),
                                      "note": "one k_fused1d launch over (2 x 25 + 1) x 34 slices: the kernel when "
                                              "the launch is large enough to fill the GPU"}
            del cb, sb
        if cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    cores = host_cores()
                    Dc = 64 if kind == 0 else (512 if kind == 2 else 2048)   # linear_d: 6144-bit MPFR, 6.4 ms / point
                    jobs = [(kind, a[k % n], 0 if eta is None else eta[k % n], Dc, sigma) for k in range(cores)]
                    t0 = time.perf_counter()
                    with mp.get_context("fork").Pool(cores) as pool:
                        res = pool.map(_ref_worker_1d, jobs)
                    w = time.perf_counter() - t0
                    sec["cpu_baseline"] = {"value": sum(res) / w, "unit": "cells/s", "cores": cores, "kind": "reference",
                                           "sample": f"{cores} slices (one per worker process) at D={Dc}, {w:.1f} s wall; "
                                                     "the cost per cell does not depend on D"}
            except Exception as exc:  # pragma: no cover
                sec["cpu_baseline"] = {"value": None, "unit": "cells/s", "cores": 0, "kind": "reference",
                                       "sample": f"failed: {exc}"}
        out[name] = sec
    return out


def _ref_worker_1d(job):
    kind, a, eta, D, sigma = job
    from oracle import ref
    d, r = synthetic_d_r(20482048)
    if kind == 2:
        P = ref.RefDiagonalParameters(M, sigma, S, d, r, eta_bound=2)
        sl = ref.diagonal_distribution_slice_compute(P, D, a, eta)
    else:
        P = ref.RefParameters(M, S, d, r, T_PARAM)
        sl = ref.linear_distribution_slice_compute(P, D, a, kind)
    return sl.cells.size


def _ref_worker_so(job):
    a, b, D = job
    from oracle import ref
    d, r = synthetic_d_r(20482048)
    P = ref.RefParameters(M, S, d, r, T_PARAM)
    return ref.distribution_slice_compute(P, D, a, b, method=1).cells.size


def sigma_optimal_section(ctx, qb, torch, stream, P, coords, peak_flops, cpu_baseline=True):
    """-sigma-optimal (probability_approx_optimal_sigma / _adjust_sigma, src/probability.cpp:20-148) on
    the whole workload at D = 128. l = 2048 is in the large-l regime: cells from the fused kernel, the
    serial walk as a prefix minimum in closed form (k_so_fast). The general fixed-point iteration
    (any l) is timed beside it on the first 64 slices."""
    import multiprocessing as mp
    a_d, a_r = [c[0] for c in coords], [c[1] for c in coords]

    def timed(ad, ar, reps):
        plan = ctx.plan2d(P, qb.DISTRIBUTION_SLICE_COMPUTE_METHOD_OPTIMAL_LOCAL_SIGMA, True, DIM, ad, ar)
        cells = torch.empty(plan.cells, dtype=torch.float64, device="cuda")
        summ = torch.empty(len(ad) * 8, dtype=torch.float64, device="cuda")
        l0 = ctx.launch_count
        plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        launches = ctx.launch_count - l0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tot = plan.cells
        tp, te, fl = plan.finish(summ.cpu().numpy())
        plan.close()
        return ms, tot, launches, tp, te

    ms, tot, launches, tp, te = timed(a_d, a_r, 5)
    os.environ["QB200_SO_FAST"] = "0"
    try:
        ms_g, tot_g, launches_g, tp_g, te_g = timed(a_d[:64], a_r[:64], 2)
    finally:
        del os.environ["QB200_SO_FAST"]
    agree = float(np.max(np.abs(((te[:64] - te_g) / te_g).astype(np.float64))))
    t0 = time.perf_counter()
    ctx.slice2d_batch(P, 1, True, DIM, a_d, a_r)
    wall = time.perf_counter() - t0
    ach = tot / (ms * 1e-3) * FLOP_PER_CELL / 1e12
    sec = {"workload": f"sigma-optimal method, all {len(coords)} slices of the workload at D={DIM} ({tot} cells)",
           "value": tot / (ms * 1e-3), "unit": "cells/s", "ms": ms, "gpu_launches": int(launches),
           "general_iteration": {"value": tot_g / (ms_g * 1e-3), "unit": "cells/s", "ms": ms_g, "slices": 64,
                                 "gpu_launches": int(launches_g),
                                 "total_error_agreement": agree},
           "roofline": {"bound": "fp64 / integer issue", "achieved": ach, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
                        "frac": ach / (peak_flops / 1e12), "flop_per_cell": FLOP_PER_CELL, "traffic": None,
                        "kernel": "k_fused2d (cells, quick-method constants) + k_so_fast (one thread per point: norm by "
                                  "angle addition, sigma* in closed form, running minimum, error sums)"},
           "e2e": {"value": tot / wall, "unit": "cells/s", "ms": wall * 1e3, "h2d_bytes_per_step": int(8 * len(coords)),
                   "d2h_bytes_per_step": int(tot * 8), "api": "qb200_slice2d_compute, method 1"}}
    if cpu_baseline:
        try:
            from oracle import ref
            if ref.available():
                cores = host_cores()
                jobs = [(coords[k % len(coords)][0], coords[k % len(coords)][1], 32) for k in range(cores)]
                t0 = time.perf_counter()
                with mp.get_context("fork").Pool(cores) as pool:
                    res = pool.map(_ref_worker_so, jobs)
                w = time.perf_counter() - t0
                sec["cpu_baseline"] = {"value": sum(res) / w, "unit": "cells/s", "cores": cores, "kind": "reference",
                                       "sample": f"{cores} slices (one per worker process) at D=32, method 1, {w:.1f} s wall"}
        except Exception as exc:  # pragma: no cover
            sec["cpu_baseline"] = {"value": None, "unit": "cells/s", "cores": 0, "kind": "reference",
                                   "sample": f"failed: {exc}"}
    return sec


def saturation_section(ctx, qb, torch, stream, timer, P, coords, rank, tp128_all, world, barrier, allmax, steps):
    """The same distribution in the generator's DEFAULT mode (dimension heuristic): every coordinate at
    its initial dimension, the upgraded ones again at 256 / 512; this rank's share of each list.
    Device resident, and end to end through the C ABI (the 512 ones scaled to 256 on the device)."""
    import ctypes as C
    dims = heuristic_dimensions(coords, tp128_all)
    everything = {}
    for i in range(len(coords)):
        init, fin = dims[i]
        everything.setdefault(init, []).append(i)
        if fin != init:
            everything.setdefault(fin, []).append(i)
    # every dimension's list is dealt out on its own (a 512 slice is 16 times a 128 slice: the
    # farm hands them out one by one, a static partition has to balance them per dimension)
    from qunundrum_b200 import shard
    lists = {D: [v[k] for k in shard.partition(len(v), max(1, world), rank)] for D, v in everything.items()}
    lists = {D: v for D, v in lists.items() if v}
    plans, bufs, cells_total = [], [], 0
    for D in sorted(lists):
        idx = lists[D]
        pl = ctx.plan2d(P, 0, True, D, [coords[i][0] for i in idx], [coords[i][1] for i in idx])
        plans.append(pl)
        bufs.append((torch.empty(max(1, pl.cells), dtype=torch.float64, device="cuda"),
                     torch.empty(max(8, len(idx) * 8), dtype=torch.float64, device="cuda")))
        cells_total += pl.cells

    # the three batches (one per dimension) are independent: each runs on a stream of its own, forked
    # from and joined to the timed stream, so that their tails and fixed costs overlap
    side = [torch.cuda.Stream() for _ in plans[1:]]

    def step():
        fork = torch.cuda.Event()
        fork.record(stream)
        for k, (pl, (c, sm)) in enumerate(zip(plans, bufs)):
            st = stream if k == 0 else side[k - 1]
            if k:
                st.wait_event(fork)
            pl.run(c.data_ptr(), sm.data_ptr(), st.cuda_stream)
        for st in side:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)
    for _ in range(3):
        step()
    barrier()
    l0 = ctx.launch_count
    ms = timer.run(step, steps)
    launches = ctx.launch_count - l0
    barrier()
    ms = allmax(ms) / steps
    all_cells = allmax(float(cells_total), "sum")
    for pl in plans:
        pl.close()
    del bufs
    # end to end: host coordinates in, stored cells out (doubles; long doubles for the scaled ones)
    L = qb.lib()
    stored, d2h = {}, 0
    for D in sorted(lists):
        n = len(lists[D])
        per = min(D, 256) ** 2
        nbytes = n * per * (16 if D > 256 else 8)
        ptr = L.qb200_host_alloc(max(16, nbytes))
        stored[D] = (ptr, n, per)
        d2h += nbytes + n * 64

    def e2e_step():
        for D in sorted(lists):
            idx = lists[D]
            ptr, n, per = stored[D]
            ad, ar = [coords[i][0] for i in idx], [coords[i][1] for i in idx]
            if D > 256:
                out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(n * per * 16,)) \
                    .view(np.longdouble).reshape(n, per)
                a_d, a_r = np.ascontiguousarray(ad, dtype=np.int32), np.ascontiguousarray(ar, dtype=np.int32)
                tp = np.zeros(n, dtype=np.longdouble)
                te = np.zeros(n, dtype=np.longdouble)
                fl = np.zeros(n, dtype=np.uint32)
                p = P._c()
                rc = L.qb200_slice2d_compute_scaled(ctx.h, C.byref(p), 0, 1, D, 256, n, a_d.ctypes.data, a_r.ctypes.data,
                                                    out.ctypes.data, tp.ctypes.data, te.ctypes.data, fl.ctypes.data)
                if rc:
                    raise SystemExit("bench.py: qb200_slice2d_compute_scaled failed")
            else:
                out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n, per))
                ctx.slice2d_batch(P, 0, True, D, ad, ar, out=out)
    e2e_step()
    barrier()
    k = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(k):
        e2e_step()
    torch.cuda.synchronize()
    wall = allmax(time.perf_counter() - t0) / k
    for ptr, _, _ in stored.values():
        L.qb200_host_free(C.c_void_p(ptr))
    d2h_all = allmax(float(d2h), "sum")
    return {"workload": ("the same distribution in -dim-heuristic mode: every coordinate at its initial dimension "
                         "(128; 256 / 512 on the diagonal tail), the upgraded ones re-computed at 256 / 512 "
                         "(src/main_generate_distribution.cpp:1222-1294), partitioned over the ranks"),
            "slices_by_dimension_this_rank": {str(D): len(v) for D, v in sorted(lists.items())},
            "cells_integrated_per_step": all_cells, "value": all_cells / (ms * 1e-3), "unit": "cells/s",
            "ms_per_step": ms, "gpu_launches_per_step": int(launches // max(1, steps)),
            "e2e": {"value": all_cells / wall, "unit": "cells/s", "ms_per_step": wall * 1e3,
                    "d2h_bytes_per_step": d2h_all,
                    "api": "qb200_slice2d_compute per dimension + qb200_slice2d_compute_scaled (512 -> 256 on the device)"}}


def run_ours(args, rank, world, local_rank):
    import torch
    import qunundrum_b200 as qb
    from qunundrum_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the slice integrators have no CPU path")
    numa = bind_near_gpu(local_rank) if world > 1 else "numa: single rank, not bound"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's banner (NCCL_DEBUG=VERSION on the GPU
        # boxes) goes to stderr instead -- stdout is pointed at stderr while the communicator
        # is created
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(v, op="max"):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    mode = scaling_mode(args, world)
    ctx = qb.Context(local_rank)
    coords = shard.enumerate_2d(M)
    n_all = len(coords)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    peak_flops = ctx.measure_fp64_peak()
    import ctypes as C
    L = qb.lib()

    def measure(P, my, tag):
        """Device-resident and end-to-end timing of this rank's slices `my` of the distribution P."""
        a_d = np.array([coords[i][0] for i in my], dtype=np.int32)
        a_r = np.array([coords[i][1] for i in my], dtype=np.int32)
        n = len(my)
        plan = ctx.plan2d(P, qb.DISTRIBUTION_SLICE_COMPUTE_METHOD_HEURISTIC_SIGMA, True, DIM, a_d, a_r)
        if plan.algorithm != 2:
            raise SystemExit("bench.py: the fused kernel was not selected")
        cells = torch.empty(max(1, plan.cells), dtype=torch.float64, device="cuda")
        summ = torch.empty(max(8, n * 8), dtype=torch.float64, device="cuda")
        flush = plan.cells * 8 < (256 << 20)
        timer = DeviceTimer(torch, stream, flush)

        def step():
            plan.run(cells.data_ptr(), summ.data_ptr(), stream.cuda_stream)
        for _ in range(max(3, args.warmup)):
            step()
        sampler = ClockSampler(local_rank) if tag == "main" else None
        barrier()
        if sampler:
            sampler.start()
        l0 = ctx.launch_count
        ms = timer.run(step, args.steps)
        barrier()
        launches = ctx.launch_count - l0
        clocks = None
        if sampler:
            # keep the clocks record meaningful for short runs: sample a little longer under load
            t_end = time.perf_counter() + 0.6
            while time.perf_counter() < t_end:
                step()
                torch.cuda.synchronize()
            clocks = sampler.stop()
        my_ms = ms / args.steps
        ms_per_step = allmax(ms) / args.steps
        h_summ = summ.cpu().numpy()[:n * 8]
        tp, te, fl = plan.finish(h_summ)
        # ---- end to end through the synchronous C ABI with host buffers ------------------
        nbytes = plan.cells * 8
        hptr = L.qb200_host_alloc(max(8, nbytes))
        if not hptr:
            raise SystemExit("bench.py: pinned host allocation failed")
        h_cells = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_double)), shape=(n, DIM * DIM))
        e2e_steps = max(1, min(args.steps, 10))
        for _ in range(2):
            ctx.slice2d_batch(P, 0, True, DIM, a_d, a_r, out=h_cells)
        barrier()
        l1 = ctx.launch_count
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, tp2, te2, fl2 = ctx.slice2d_batch(P, 0, True, DIM, a_d, a_r, out=h_cells)
        torch.cuda.synchronize()
        my_wall = (time.perf_counter() - t0) / e2e_steps
        e2e_launches = ctx.launch_count - l1
        wall = allmax(my_wall)
        mass = float(tp.sum())
        if abs(float(tp2.sum()) - mass) > 1e-13 or abs(h_cells.sum() - mass) > 1e-9:
            raise SystemExit("bench.py: end-to-end results differ from the device-resident run")
        return dict(plan=plan, cells=cells, summ=summ, h_summ=h_summ, tp=tp, tp2=tp2, ms_per_step=ms_per_step,
                    my_ms=my_ms, wall=wall, my_wall=my_wall, launches=launches, e2e_launches=e2e_launches,
                    e2e_steps=e2e_steps, clocks=clocks, n=n, nbytes=nbytes, hptr=hptr, h_cells=h_cells, mass=mass,
                    timer=timer, flush=flush,
                    h2d=int(a_d.nbytes + a_r.nbytes + 2 * ((M + 7) // 8) + n * 40), d2h=int(nbytes + n * 64))

    # ---- the headline measurement ------------------------------------------------------------
    if mode == "strong":
        d, r = synthetic_d_r(20482048)                 # ONE distribution, partitioned
        my = shard.partition(n_all, world, rank)
    else:
        d, r = synthetic_d_r(20482048 + rank)          # one distribution per rank
        my = np.arange(n_all)
    P = qb.Parameters(M, S, d, r, T_PARAM)
    R = measure(P, my, "main")
    cells_job = n_all * DIM * DIM * (1 if mode == "strong" else world)
    value = cells_job / (R["ms_per_step"] * 1e-3)
    e2e_value = cells_job / R["wall"]
    # results sanity (and the only cross-rank traffic): gather the slice summaries
    if mode == "strong":
        table = shard.gather_summaries(my, R["h_summ"].reshape(-1, 8), n_all)
        total_mass = allmax(R["mass"], "sum")
    else:
        table = shard.gather_summaries(rank * n_all + np.arange(n_all), R["h_summ"].reshape(-1, 8), n_all * world)
        total_mass = R["mass"]
    if not (0.4999 < total_mass < 0.5):
        raise SystemExit(f"bench.py: captured mass {total_mass} is wrong")
    per_rank = None
    if dist is not None:
        t = torch.tensor([R["my_ms"], R["my_wall"] * 1e3, float(R["n"])], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        per_rank = {"device_ms_per_step": [float(x[0]) for x in g], "e2e_ms_per_step": [float(x[1]) for x in g],
                    "slices": [int(x[2]) for x in g]}

    # ---- the other scaling mode, for N > 1 (reported under `weak` / `strong`) -----------------
    other = None
    if world > 1 and not args.no_other_scaling:
        L.qb200_host_free(C.c_void_p(R["hptr"]))
        R["hptr"] = None
        R["plan"].close()
        if mode == "strong":
            P2 = qb.Parameters(M, S, *synthetic_d_r(20482048 + rank), T_PARAM)
            O = measure(P2, np.arange(n_all), "other")
            cj = n_all * DIM * DIM * world
            name = "weak"
        else:
            P2 = qb.Parameters(M, S, *synthetic_d_r(20482048), T_PARAM)
            O = measure(P2, shard.partition(n_all, world, rank), "other")
            cj = n_all * DIM * DIM
            name = "strong"
        other = {"scaling": name, "value": cj / (O["ms_per_step"] * 1e-3), "unit": "cells/s",
                 "ms_per_step": O["ms_per_step"],
                 "e2e": {"value": cj / O["wall"], "unit": "cells/s", "ms_per_step": O["wall"] * 1e3,
                         "d2h_bytes_per_step": int(O["d2h"])},
                 "note": ("one whole distribution per rank (different d, r): N replicas" if name == "weak" else
                          "ONE distribution partitioned over the ranks")}
        L.qb200_host_free(C.c_void_p(O["hptr"]))
        O["plan"].close()

    # ---- saturation workload: the dimension-heuristic distribution, partitioned ----------------
    saturation = None
    if not args.no_saturation:
        Ps = qb.Parameters(M, S, *synthetic_d_r(20482048), T_PARAM)
        # the totals at D = 128 decide which coordinates the heuristic upgrades
        _, tp_all, _, _ = ctx.slice2d_batch(Ps, 0, True, DIM, [c[0] for c in coords], [c[1] for c in coords])
        saturation = saturation_section(ctx, qb, torch, stream, DeviceTimer(torch, stream, False), Ps, coords, rank,
                                        [float(x) for x in tp_all], world, barrier, allmax, max(3, min(args.steps, 10)))

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            import multiprocessing as mp
            from oracle import ref
            if ref.available():
                cores = host_cores()
                with mp.get_context("fork").Pool(cores) as pool:
                    c, w = reference_cpu_pass(pool, cores, coords, 1)
                cpu = {"value": c / w, "unit": "cells/s", "cores": cores, "kind": "reference",
                       "sample": (f"{cores} slices (one per worker process; the first {cores} of the "
                                  f"enumerator order), D=128 Richardson, 192-bit MPFR, {w:.1f} s wall"),
                       "cells_per_s_per_core": c / w / cores}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "unit": "cells/s", "cores": 0, "kind": "reference",
                   "sample": f"failed: {exc}"}

    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"

    sections = {}
    if world == 1 and not args.no_sections:
        sections.update(one_dimensional_sections(ctx, qb, torch, stream, peak_flops, hbm_peak,
                                                 cpu_baseline=not args.no_cpu_baseline))
        sections["sigma_optimal"] = sigma_optimal_section(ctx, qb, torch, stream, P, coords, peak_flops,
                                                          cpu_baseline=not args.no_cpu_baseline)
    tau = None
    if world == 1 and not args.no_tau:
        tau = tau_section(ctx, qb, torch, stream, R["h_cells"], R["tp2"], coords, hbm_peak,
                          cpu_baseline=not args.no_cpu_baseline)
    diagk = None
    if world == 1 and not args.no_tau:
        diagk = diagk_section(ctx, qb, torch, stream, cpu_baseline=not args.no_cpu_baseline)
    exact = None
    if world == 1 and not args.no_tau:
        exact = exact_section(ctx, qb, torch, stream, cpu_baseline=not args.no_cpu_baseline)
    if R["hptr"]:
        L.qb200_host_free(C.c_void_p(R["hptr"]))
    if dist is not None:
        dist.barrier()
    text = None
    if world == 1 and not args.no_text:
        R["plan"].run(R["cells"].data_ptr(), R["summ"].data_ptr(), stream.cuda_stream)   # the step's own cells
        torch.cuda.synchronize()
        text = text_section(ctx, qb, torch, stream, R["cells"], hbm_peak)

    if rank == 0:
        prof = profile_constants() or {}
        my_cells = R["n"] * DIM * DIM
        cells_per_s_gpu = my_cells / (R["my_ms"] * 1e-3)        # this GPU
        canonical = cells_per_s_gpu * FLOP_PER_CELL / 1e12
        peak = peak_flops / 1e12
        # executed: FP64 instructions per cell of the three class kernels (ncu, profiles/fused2d_latest.json,
        # weighted by the classes' share of the workload) x cells/s over the DFMA-instruction rate of the
        # same GPU in the same run (the microkernel's flop/s / 2)
        inst = prof.get("fp64_inst_per_cell")
        executed = None if not inst else cells_per_s_gpu * inst / (peak_flops / 2.0)
        out_gbs = R["nbytes"] / (R["my_ms"] * 1e-3) / 1e9
        compact = {}
        for k, v in sections.items():
            compact[k] = {"cells_per_s": v["value"], "ms": v["ms"], "frac": v["roofline"]["frac"],
                          "e2e_cells_per_s": v["e2e"]["value"],
                          "cpu_cells_per_s": (v.get("cpu_baseline") or {}).get("value")}
            if "at_eta_bound_25" in v:   # the same kernel with a launch that fills the GPU
                compact[k]["full_gpu_cells_per_s"] = v["at_eta_bound_25"]["value"]
                compact[k]["full_gpu_frac"] = v["at_eta_bound_25"]["frac"]
        if text:
            compact["text_export"] = {"values_per_s": text["export"]["values_per_s"],
                                      "hbm_frac": text["export"]["roofline"]["frac"]}
            compact["text_import"] = {"values_per_s": text["import"]["values_per_s"],
                                      "hbm_frac": text["import"]["roofline"]["frac"]}
        if tau:
            compact["tau"] = {"samples_per_s": tau["value"], "hbm_frac": tau["roofline"]["frac"],
                              "e2e_samples_per_s": tau["e2e"]["value"],
                              "cpu_samples_per_s": (tau.get("cpu_baseline") or {}).get("value")}
        if diagk:
            compact["diagk"] = {"samples_per_s": diagk["value"], "e2e_samples_per_s": diagk["e2e"]["value"],
                                "cpu_samples_per_s": (diagk.get("cpu_baseline") or {}).get("value")}
        if exact:
            compact["exact"] = {"samples_per_s": exact["value"], "e2e_samples_per_s": exact["e2e"]["value"],
                                "cpu_samples_per_s": (exact.get("cpu_baseline") or {}).get("value")}
        if saturation:
            compact["saturation"] = {"cells_per_s": saturation["value"], "ms": saturation["ms_per_step"],
                                     "e2e_cells_per_s": saturation["e2e"]["value"]}
        if other:
            compact[other["scaling"]] = {"cells_per_s": other["value"], "e2e_cells_per_s": other["e2e"]["value"]}
        line = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": R["ms_per_step"],
            "higher_is_better": True, "scaling": mode, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, mode),
            "roofline": {
                "bound": "fp64", "achieved": None if executed is None else executed * peak, "peak": peak,
                "unit": "TFLOP/s", "frac": executed, "frac_canonical": canonical / peak,
                "achieved_canonical": canonical, "traffic": prof.get("dram_bytes_per_launch"),
                "kernel": "k_fused2d<MODE 0, class 0 / 1 / 2>",
                "flop_per_cell_canonical": FLOP_PER_CELL,
                "fp64_inst_per_cell": inst,
                "fp64_pipe_active_frac_ncu": prof.get("fp64_pipe_active_frac"),
                "profile_matches_source": prof.get("kernel_source_sha") == kernel_source_sha(),
                "peak_source": "DFMA microkernel measured in this run on this GPU "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "hbm": {"algorithmic_gbs": out_gbs, "peak_gbs": hbm_peak, "frac": out_gbs / hbm_peak,
                        "peak_source": hbm_src},
                "sections": compact,
                "note": "rank 0's GPU; step time includes the three small table/summary kernels (<1.5%)",
                "frac_note": ("frac = executed FP64 instructions (fp64_inst_per_cell from the committed ncu capture "
                              "of this kernel source x cells/s of this run) over the DFMA-instruction rate measured "
                              "in this run; frac_canonical uses SURVEY 8(d)'s 1600 flop per cell, which the kernel "
                              "undercuts (angle addition from per-row / per-column tables), so it exceeds 1"),
            },
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": R["h2d"],
                    "d2h_bytes_per_step": R["d2h"], "steps": R["e2e_steps"],
                    "ms_per_step": R["wall"] * 1e3, "bytes_are": "per rank",
                    "api": "qb200_slice2d_compute (synchronous C ABI, pinned host result buffer)",
                    "host_placement": numa},
            "gpu_launches": int(R["launches"]),
            "e2e_gpu_launches": int(R["e2e_launches"]),
            "clocks": R["clocks"],
            "captured_mass_per_distribution": total_mass,
            "gathered_summaries": None if table is None else int(table.shape[0]),
            "per_rank": per_rank,
            "other_scaling": other,
            "saturation": saturation,
            "sections": sections or None,
            "text": text,
            "tau": tau,
            "diagk": diagk,
            "exact": exact,
        }
        print(json.dumps(line))
    try:
        R["plan"].close()
    except Exception:
        pass
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text", action="store_true")
    ap.add_argument("--no-tau", action="store_true")
    ap.add_argument("--no-sections", action="store_true")
    ap.add_argument("--no-saturation", action="store_true")
    ap.add_argument("--no-other-scaling", action="store_true")
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"],
                    help="N > 1: strong = ONE distribution partitioned over the ranks (default), "
                         "weak = one distribution per rank")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        if args.no_cpu_baseline:
            cmd.append("--no-cpu-baseline")
        if args.no_text:
            cmd.append("--no-text")
        if args.no_tau:
            cmd.append("--no-tau")
        for flag, on in (("--no-sections", args.no_sections), ("--no-saturation", args.no_saturation),
                         ("--no-other-scaling", args.no_other_scaling)):
            if on:
                cmd.append(flag)
        cmd += ["--scaling", args.scaling]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
