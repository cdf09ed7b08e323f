"""profiles/fused2d_latest.json (what bench.py quotes as roofline.traffic / executed_frac) from
the ncu summary of one step of the bench workload:

    python tests/tools/ncu_summary.py gpurun_out/fused2d.ncu-rep profiles/rNN_fused2d_ncu_full
    python tests/tools/fused2d_latest.py profiles/rNN_fused2d_ncu_full.json
"""
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
out = {
    "source": f"{os.path.relpath(src, ROOT)} (ncu --set full --clock-control none --import-source on, "
              "tests/tools/prof_t2d.py: one step of the bench workload)",
    "kernels": [name(r) for r in ks],
    "dram_bytes_per_launch": sum(r["dram_read"] + r["dram_write"] for r in ks),
    "fp64_pipe_active_frac": sum(r["fp64_pipe_active_pct"] * r["duration"] for r in ks) / dur / 100.0,
    "kernel_time_under_ncu_s": dur,
    "per_kernel": [{"kernel": name(r), "duration_s": r["duration"], "fp64_pipe_active_pct": r["fp64_pipe_active_pct"],
                    "dram_bytes": r["dram_read"] + r["dram_write"], "registers": r["registers"]} for r in ks],
    "note": "one step = three launches of k_fused2d (one per slice class); traffic and pipe share are summed / "
            "duration-weighted over them",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "fused2d_latest.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
