for extra in "--no-graph" ""; do
for w in 0 1; do
FBKST_ATTN_WIDE=$w timeout 300 python bench.py --config cfg1 --steps 10 --warmup 3 $extra 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = r.get('parity') or {}
print('cfg1 wide=$w $extra value %.2fM ms/step %.3f parity %s %s lengths_equal %s' % (r['value']/1e6, r['ms_per_step'], p.get('max_rel'), p.get('elementwise'), p.get('lengths_equal')))
"
done; done
FBKST_ATTN_WIDE=0 timeout 300 python bench.py --config cfg3 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = r.get('parity') or {}
print('cfg3 wide=0 parity', p)
"
