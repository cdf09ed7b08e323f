#!/bin/bash
# Round 2, fourth GPU pass (1 GPU): class-overlap A/B, guide-table sampler A/B, full gpu test suite.
set -x
mkdir -p gpurun_out
for v in 1 0; do
  QB200_OVERLAP_CLASSES=$v timeout 300 python tests/tools/prof_t2d.py 20 128 > gpurun_out/c4_t2d_overlap$v.txt 2>&1
  QB200_OVERLAP_CLASSES=$v timeout 300 python tests/tools/prof_t2d.py 20 256 >> gpurun_out/c4_t2d_overlap$v.txt 2>&1
done
cat gpurun_out/c4_t2d_overlap1.txt gpurun_out/c4_t2d_overlap0.txt
QB200_SAMPLER_GUIDE=1 timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c4_sampler_guide.json 2> gpurun_out/c4_sampler_guide.err
QB200_SAMPLER_GUIDE=0 timeout 300 python tests/tools/prof_sampler.py > gpurun_out/c4_sampler_plain.json 2> gpurun_out/c4_sampler_plain.err
python - <<'PY'
import json
for f in ("c4_sampler_guide", "c4_sampler_plain"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "%.4g samples/s" % d["value"], d["ms"], "e2e %.4g" % d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
for v in 0 3 4 6; do
  QB200_DIAGK_CTAS_PER_SM=$v timeout 300 python tests/tools/prof_diagk.py > gpurun_out/c4_diagk_ctas$v.json 2> gpurun_out/c4_diagk_ctas$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/c4_diagk_ctas$v.json').read().strip().splitlines()[-1]); print('diagk ctas/SM $v', d['value'], d['ms'], d['default_delta_bound']['ms'])"
done
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/c4_tests_all.txt 2>&1
tail -5 gpurun_out/c4_tests_all.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_sample$' -c 1 -o gpurun_out/c4_sampler python tests/tools/prof_sampler.py > gpurun_out/c4_ncu_sampler.log 2>&1
python tests/tools/ncu_summary.py gpurun_out/c4_sampler.ncu-rep gpurun_out/r02_sampler_ncu_full > /dev/null 2>&1
