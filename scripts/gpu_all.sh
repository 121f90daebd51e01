#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests.sh 2>&1 | tee gpurun_out/tests_summary.txt
timeout 300 python scripts/bench_small.py 10 ctc 2>&1 | tee gpurun_out/bench_small.txt
bash -c 'timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"'
bash -c 'timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_quick_nograph.json 2> gpurun_out/bench_quick_nograph.err; echo "bench exit=$?"'
python - <<'PY'
import json
for f in ("bench_quick", "bench_quick_nograph"):
    try:
        r = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.3fM e2e %.3fM ms/step %.4f launches %d" % (r["value"]/1e6, r["e2e"]["value"]/1e6, r["ms_per_step"], r["gpu_launches"]))
        print(" roofline", r["roofline"]["kernel"], r["roofline"]["achieved"], r["roofline"]["frac"])
        if f == "bench_quick":
            for k, v in r["kernels"].items():
                print("  %-50s %8.4f ms  n=%-3d tf=%s gbs=%s" % (k, v["ms_per_step"], v["launches_per_step"], v["tflops"], v["gbs"]))
    except Exception as e:
        print("no bench json", f, e); print(open("gpurun_out/%s.err" % f).read()[-3000:])
PY
