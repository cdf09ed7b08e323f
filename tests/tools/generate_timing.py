"""Wall clock and phases of the reference's own generate_distribution (gpu flavour: integrator,
text and collapse drop-ins) on a FULL m = 2048 two-dimensional distribution (GPU box).

    python tests/tools/generate_timing.py [--clients N] [--dim 256|0] [--prefetch 0|1] [--tag NAME]

--dim 0 = the generator's default dimension heuristic (BASELINE configs[1] as written).
Phases from the server's own stdout lines; QB200_DROPIN_STATS=1 gives the time inside the drop-ins
and the C ABI per process. Appends to gpurun_out/generate_timing.json."""
import argparse, json, os, random, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = os.path.join(ROOT, "integration", "_build")
ap = argparse.ArgumentParser()
ap.add_argument("--clients", type=int, default=1)
ap.add_argument("--dim", type=int, default=256)
ap.add_argument("--prefetch", type=int, default=1)
ap.add_argument("--tag", default="")
ap.add_argument("--keep", action="store_true")
args = ap.parse_args()
rnd = random.Random(20482048); m = 2048
r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1); d = r // 2 + rnd.randrange(r // 2)
t = tempfile.mkdtemp(); os.makedirs(t + "/distributions")
cmd = [B + "/minimpirun", "-np", str(args.clients + 1), B + "/gpu/generate_distribution", "-exp", str(d), str(r),
       *(["-dim", str(args.dim)] if args.dim else []), "2048", "1"]
env = dict(os.environ, QB200_DROPIN_STATS="1", QB200_PREFETCH=str(args.prefetch))
rep = {"tag": args.tag, "clients": args.clients, "dim": args.dim or "heuristic", "prefetch": args.prefetch}
t0 = time.time()
err = open(os.path.join(t, "stderr.txt"), "w")
p = subprocess.Popen(["stdbuf", "-oL"] + cmd, cwd=t, stdout=subprocess.PIPE, stderr=err, text=True, env=env)
marks, dims = {}, {}
for line in p.stdout:
    for key in ("Processing slice: 1 /", "Processing slice: 100 /", "Stopping node", "Sorting the slices",
                "Exporting distribution information", "Exporting collapsed distribution", "Exporting the distribution to",
                "Finished exporting"):
        if key in line and key not in marks:
            marks[key] = round(time.time() - t0, 3)
    if line.startswith("Slice dimension is:"):
        k = line.split(":")[1].strip()
        dims[k] = dims.get(k, 0) + 1
p.wait()
rep["returncode"] = p.returncode
rep["generate_wall_s"] = round(time.time() - t0, 3)
rep["marks_s"] = marks
rep["received_slices_by_dimension"] = dims
rep["files"] = {f: os.path.getsize(os.path.join(t, "distributions", f)) for f in sorted(os.listdir(t + "/distributions"))}
err.close()
rep["stats"] = [l.strip() for l in open(os.path.join(t, "stderr.txt")) if "qunundrum_b200" in l]
inside = 0.0
for l in rep["stats"]:
    import re
    mm = re.search(r"([0-9.]+) s inside the drop-in functions", l)
    if mm:
        inside += float(mm.group(1))
    for key in (r"uploads \(([0-9.]+) s\)", r"collapses \(([0-9.]+) s\)", r"export batches \(([0-9.]+) s\)"):
        mm = re.search(key, l)
        if mm:
            inside += float(mm.group(1))
rep["seconds_inside_dropins_all_processes"] = round(inside, 3)
print(json.dumps(rep, indent=1))
out = os.path.join(ROOT, "gpurun_out", "generate_timing.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
allr = json.load(open(out)) if os.path.exists(out) else []
allr.append(rep)
json.dump(allr, open(out, "w"), indent=1)
if not args.keep:
    shutil.rmtree(t, ignore_errors=True)
sys.exit(p.returncode)
