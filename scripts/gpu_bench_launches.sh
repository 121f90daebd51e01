#!/bin/bash
# Bench line + ncu launch list of three inference steps with its per-kernel summary
TAG=${1:-r02y}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
TAG=$TAG python - <<'PY'
import json, os
TAG = os.environ['TAG']
try:
    r = json.load(open("gpurun_out/%s_bench.json" % TAG))
    print("value %.3fM e2e %.3fM ms/step %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["gpu_launches"]))
    print(" roofline", r["roofline"]["achieved"], r["roofline"]["frac"], "sustained", r["roofline"]["sustained"]["value"], r["roofline"]["sustained"]["clocks"]["sm_mhz"])
    for k, v in r["kernels"].items():
        print("  %-50s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k[:50], v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
except Exception as e:
    print("no bench json", e); print(open("gpurun_out/%s_bench.err" % TAG).read()[-3000:])
PY
FBKST_PROFILE_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
  --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list exit=$? lines=$(wc -l < gpurun_out/${TAG}_launches.csv)"
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1; head -n 24 gpurun_out/${TAG}_launches.md
