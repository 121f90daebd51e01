#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N tag
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
run() { # name, args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  echo "$name N=$N rc=$?"; tail -c 600 gpurun_out/${TAG}_${name}_${N}gpu.err | grep -v "Warning\|warn" | tail -5
  python - <<PY
import json
try:
    r = json.load(open("gpurun_out/${TAG}_${name}_${N}gpu.json"))
    print("  value %.3fM  ms/step %.4f" % (r["value"]/1e6, r["ms_per_step"]), {k: r.get(k) for k in ("rank_time_ms", "step_imbalance_padded_frames", "nonfinite_outputs", "train_step_split", "allreduce") if r.get(k)})
    if r.get("e2e"): print("  e2e %.3fM" % (r["e2e"]["value"]/1e6))
except Exception as e:
    print("  no json", e)
PY
}
ONLY=${3:-all}
if [ "$ONLY" = "all" ] || [ "$ONLY" = "cfg3r" ]; then run cfg3r --config cfg3r --steps 20 --warmup 3; fi
if [ "$ONLY" = "all" ] || [ "$ONLY" = "cfg4" ]; then run cfg4 --config cfg4 --steps 10 --warmup 3; fi
if [ "$ONLY" = "all" ] || [ "$ONLY" = "cfg2" ]; then run cfg2 --steps 20 --warmup 5 --no-cpu-baseline; fi
