run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-mf --no-hybrid --no-pipeline --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3), d['plan']['t1'], d['plan']['t2'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})
"; }
for t in 12 15 16 20 25; do D4W_T1=$t run "T1=$t"; done
