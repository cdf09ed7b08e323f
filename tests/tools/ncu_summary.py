"""Summarise an .ncu-rep (raw page) into a small JSON + text table for profiles/.

    python tests/tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "launch__grid_size": "grid",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__waves_per_multiprocessor": "waves",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__maximum_warps_avg_per_active_cycle": "max_warps_per_sm",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_inst_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__inst_executed_pipe_fp64.sum": "fp64_warp_instructions",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_inst",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_inst",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__block_size": "block",
}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9,
        "s": 1.0, "Ghz": 1e9, "Mhz": 1e6, "cycle": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS and v != "":
                x = float(v.replace(",", ""))
                if KEYS[h] in ("duration", "dram_read", "dram_write", "sm_clock"):
                    x *= UNIT.get(u, 1.0)
                d[KEYS[h]] = x
        kernels.append(d)
    json.dump(kernels, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for d in kernels:
            f.write(d["kernel"] + "\n")
            for k, v in d.items():
                if k != "kernel":
                    f.write(f"  {k:28s} {v:.6g}\n")
            f.write("\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
