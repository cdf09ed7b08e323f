"""profiles/fused2d_latest.json (what bench.py quotes as roofline.traffic / executed_frac) from
the ncu summary of one step of the bench workload:

    python tests/tools/ncu_summary.py gpurun_out/fused2d.ncu-rep profiles/rNN_fused2d_ncu_full
    python tests/tools/fused2d_latest.py profiles/rNN_fused2d_ncu_full.json

The capture must carry smsp__inst_executed_pipe_fp64.sum (ncu --set full --metrics
smsp__inst_executed_pipe_fp64.sum ...): fp64_inst_per_cell = 32 x that count, summed over the three
class launches of one step, over the step's 55,083,008 cells -- the executed FP64 instructions per
output cell of THIS kernel source (kernel_source_sha), which bench.py turns into roofline.frac with
the step time and DFMA rate it measures itself.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = sys.argv[1]
rows = json.load(open(src))
rows = rows if isinstance(rows, list) else rows.get("kernels", rows)
ks = [r for r in rows if "k_fused2d" in r.get("kernel", r.get("name", ""))][:3]
name = lambda r: r.get("kernel", r.get("name"))  # noqa: E731
dur = sum(r["duration"] for r in ks)
CELLS_PER_STEP = 3362 * 128 * 128
fp64 = sum(r.get("fp64_warp_instructions", 0) for r in ks)
sha = hashlib.sha256()
for f in ("kernels_fused2d.cuh", "qmath.cuh", "integrands.cuh", "slice_cells.cuh"):
    sha.update(open(os.path.join(ROOT, "qunundrum_b200", "csrc", f), "rb").read())
out = {
    "fp64_inst_per_cell": (32.0 * fp64 / CELLS_PER_STEP) if fp64 else None,
    "fp64_warp_instructions_per_step": fp64 or None,
    "kernel_source_sha": sha.hexdigest()[:16],
    "source": f"{os.path.relpath(src, ROOT)} (ncu --set full --clock-control none --import-source on, "
              "tests/tools/prof_t2d.py: one step of the bench workload)",
    "kernels": [name(r) for r in ks],
    "dram_bytes_per_launch": sum(r["dram_read"] + r["dram_write"] for r in ks),
    "fp64_pipe_active_frac": sum(r["fp64_pipe_active_pct"] * r["duration"] for r in ks) / dur / 100.0,
    "kernel_time_under_ncu_s": dur,
    "per_kernel": [{"kernel": name(r), "duration_s": r["duration"], "fp64_pipe_active_pct": r["fp64_pipe_active_pct"],
                    "dram_bytes": r["dram_read"] + r["dram_write"], "registers": r["registers"],
                    "fp64_warp_instructions": r.get("fp64_warp_instructions")} for r in ks],
    "note": "one step = three launches of k_fused2d (one per slice class); traffic and pipe share are summed / "
            "duration-weighted over them",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "fused2d_latest.json"), "w"), indent=1)
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):   # on a GPU box: bring it home too
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fused2d_latest.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
