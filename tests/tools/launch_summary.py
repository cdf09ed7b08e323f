"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:

    python tests/tools/launch_summary.py gpurun_out/bench_launches.csv > profiles/rNN_bench_launches_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
    rows.append((name, us))
tot = defaultdict(float)
cnt = defaultdict(int)
for n, us in rows:
    tot[n] += us
    cnt[n] += 1
total = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu-baseline")
print("(per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes; k_dfma_peak is the FP64")
print("peak microbenchmark, the k_text_* launches are the `text` section and k_seg_build / k_sample / k_tau_reduce the")
print("`tau` section, k_diagk* the `diagk` section: all outside the timed step)\n")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
for n in sorted(tot, key=lambda k: -tot[k]):
    print(f"{n[:60]:60s} {cnt[n]:8d} {tot[n]:10.1f} {100 * tot[n] / total:6.2f}%")
step = {n: t for n, t in tot.items() if re.search(r"k_fused|k_axis2d", n)}
st = sum(step.values())
print("\nshares within the integration step (fused + table + summary kernels):")
for n in sorted(step, key=lambda k: -step[k]):
    print(f"  {n[:56]:56s} {100 * step[n] / st:6.2f}%")
