#!/bin/bash
# A/B of the attention kernels (FBKST_ATTN_WIDE=0 / 1) through bench.py on the configurations given as arguments
for cfg in "$@"; do
  for w in 0 1 0 1; do
    FBKST_ATTN_WIDE=$w timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = r.get('kernels', {}).get('attention', {})
p = r.get('parity') or {}
print('$cfg wide=$w value %.2fM ms/step %.3f e2e %.2fM attention %.4f ms parity %s %s lengths_equal %s clocks %s' % (r['value']/1e6, r['ms_per_step'], r['e2e']['value']/1e6, k.get('ms_per_step', -1), p.get('max_rel'), p.get('elementwise'), p.get('lengths_equal'), r.get('clocks')))
"
  done
done
