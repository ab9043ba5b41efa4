export D4W_COL_PIPE3=1
python -m pytest tests/test_fullsize_gpu.py::test_config2_fan_mask_direct tests/test_fk_gpu.py -m gpu -q -s --timeout 900 -p no:cacheprovider 2>&1 | grep -E "config 2|passed|failed" 
run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-mf --no-hybrid --no-pipeline --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],3), d['plan']['col_scheme'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})
"; }
D4W_PIPE3_NP=4 D4W_PIPE3_OCC=3 run "np4 occ3"
D4W_PIPE3_NP=2 D4W_PIPE3_OCC=4 run "np2 occ4"
D4W_PIPE3_NP=2 D4W_PIPE3_OCC=3 run "np2 occ3"
D4W_PIPE3_NP=4 D4W_PIPE3_OCC=2 run "np4 occ2"
D4W_PIPE3_NP=8 D4W_PIPE3_OCC=2 run "np8 occ2"
D4W_PIPE3_NP=4 D4W_PIPE3_OCC=3 D4W_PIPE3_CQ=100 run "np4 occ3 cq100"
D4W_PIPE3_NP=4 D4W_PIPE3_OCC=3 D4W_PIPE3_CQ=25 run "np4 occ3 cq25"
D4W_PIPE3_NP=2 D4W_PIPE3_OCC=4 D4W_PIPE_LAG=3 run "np2 occ4 lag3"
