"""Per-source-line instruction counts of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on):  python tests/tools/ncu_source_lines.py rep.ncu-rep <kernel regex> [top N]"""
import csv, io, re, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
lines = raw.splitlines()
i = 0
files = {}
done = set()
cur_file = cur_fn = None
rows = []
for row in csv.reader(lines):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1]; continue
    if row[0] == "Function Name" or row[0] == "Kernel Name":
        cur_fn = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if row[0] == "Address":
        hdr = None; continue
    if hdr and row[0].isdigit() and re.search(pat, cur_fn or ""):
        # source text may hold quotes / commas that break the CSV: index from the right
        ki = len(hdr) - hdr.index("Instructions Executed")
        ks = len(hdr) - hdr.index("# Samples")
        try:
            rows.append((cur_file, int(row[0]), row[1], int(row[-ki] or 0), int(row[-ks] or 0), cur_fn))
        except ValueError:
            pass
# one launch only: keep rows of the first occurrence of each (file, line)
seen, uniq = set(), []
for r in rows:
    k = (r[0], r[1])
    if k in seen:
        continue
    seen.add(k); uniq.append(r)
total = sum(r[3] for r in uniq) or 1
samples = sum(r[4] for r in uniq) or 1
print(f"total warp instructions {total}, samples {samples}")
for r in sorted(uniq, key=lambda r: -r[3])[:top]:
    print(f"{100*r[3]/total:5.1f}% inst {100*r[4]/samples:5.1f}% smp  {r[0].split('/')[-1]}:{r[1]:<4d} {r[2].strip()[:90]}")
