for cfg in "D4W_XCORR_FUSED=1" "D4W_XCORR_FUSED=0" "D4W_BLOCK_PLAN=4,25,25" "D4W_BLOCK_PLAN=10,10,25" "D4W_BLOCK_PLAN=25,20,5" "D4W_BLOCK_PLAN=20,5,25"; do env $cfg timeout 200 python scripts/gpu_bench_xcorr.py 2>&1 | tail -1; done
timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 300 python scripts/gpu_bench_rows.py 2>&1 | tail -14
timeout 600 python bench.py > gpurun_out/r01h_bench_n1.json 2> gpurun_out/r01h_bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r01h_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['fk_plus_matched_filter']['value'], d['fk_plus_matched_filter']['ms_per_step'])"
