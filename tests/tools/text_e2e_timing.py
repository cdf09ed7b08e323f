"""Wall-clock of the reference's own executables on a FULL m = 2048 two-dimensional
distribution, with and without the drop-ins (GPU box):
  generate_distribution (gpu flavour: integrators + text drop-in)     phases
  filter_distribution   gpu flavour vs ref flavour on that 3.1 GB file  (import + export)
and cmp of the two filtered files. Writes gpurun_out/text_e2e_timing.json."""
import json, os, random, subprocess, sys, tempfile, time, filecmp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = os.path.join(ROOT, "integration", "_build")
clients = int(sys.argv[1]) if len(sys.argv) > 1 else 1
run_ref = "--no-ref" not in sys.argv
pin = "--pin" in sys.argv        # one visible GPU per rank (integration/pin_gpu.sh)
rnd = random.Random(20482048); m = 2048
r = 2 ** (m - 1) + 1 + rnd.randrange(2 ** (m - 1) - 1); d = r // 2 + rnd.randrange(r // 2)
t = tempfile.mkdtemp(); os.makedirs(t + "/distributions")
cmd = [B + "/minimpirun", "-np", str(clients + 1),
       *([os.path.join(ROOT, "integration", "pin_gpu.sh")] if pin else []),
       B + "/gpu/generate_distribution", "-exp", str(d), str(r), "-dim", "256", "2048", "1"]
rep = {"clients": clients, "one_visible_gpu_per_rank": pin}
t0 = time.time()
p = subprocess.Popen(["stdbuf", "-oL"] + cmd, cwd=t, stdout=subprocess.PIPE, text=True)
marks = {}
for line in p.stdout:
    for key in ("Processing slice: 1 /", "Processing slice: 100 /", "Stopping node", "Sorting the slices", "Exporting distribution information",
                "Exporting the distribution to", "Finished exporting"):
        if key in line and key not in marks:
            marks[key] = round(time.time() - t0, 2)
p.wait()
rep["generate_wall_s"] = round(time.time() - t0, 2)
rep["generate_marks_s"] = marks
main = [f for f in os.listdir(t + "/distributions") if f.startswith("distribution-") and f.endswith(".txt")][0]
path = os.path.join(t, "distributions", main)
rep["file_bytes"] = os.path.getsize(path)
for flavour in (["gpu", "ref"] if run_ref else ["gpu"]):
    w = os.path.join(t, flavour); os.makedirs(w + "/distributions")
    t0 = time.time()
    q = subprocess.run([os.path.join(B, flavour, "filter_distribution"), path], cwd=w, capture_output=True, text=True)
    rep[f"filter_{flavour}_wall_s"] = round(time.time() - t0, 2)
    assert q.returncode == 0, q.stdout + q.stderr
if run_ref:
    a = os.path.join(t, "gpu", "distributions", "filtered-" + main)
    b = os.path.join(t, "ref", "distributions", "filtered-" + main)
    rep["filtered_bytes"] = os.path.getsize(a)
    rep["filtered_files_identical"] = filecmp.cmp(a, b, shallow=False)
print(json.dumps(rep, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "text_e2e_timing.json"), "w"), indent=1)
