#!/bin/bash
# Long soaks of the pipelined inference path (default attention kernel) on three configurations: no hang (the
# mbarrier watchdogs trap), outputs finite, parity of the sample taken AFTER the soak
mkdir -p gpurun_out
for spec in "cfg2 15" "cfg3 10" "cfg1 5" "cfg5 5"; do
  set -- $spec
  timeout 600 python bench.py --config $1 --steps 10 --warmup 3 --soak-seconds $2 2>gpurun_out/soak_$1.err | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = r.get('parity') or {}; s = r['roofline']['sustained']
print('$1 soak %d steps %.1f s: %.2f M frames/s at %s MHz %s | parity after the soak %s / %s non-finite %s lengths_equal %s' % (
    s['steps'], s['seconds'], s['value'] / 1e6, s['clocks']['sm_mhz'], s['clocks']['reasons'], p.get('max_rel'), p.get('elementwise'), p.get('nonfinite_ours'), p.get('lengths_equal')))
" || { echo "$1 FAILED"; tail -5 gpurun_out/soak_$1.err; }
done
