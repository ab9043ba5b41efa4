timeout 300 python -u -m pytest tests/test_fk_gpu.py -x -q --timeout 200 2>&1 | tail -3
for cfg in "D4W_ROW_FUSED=1" "D4W_ROW_FUSED=0" "D4W_ROW_FUSED=1 D4W_ROW_THREADS=512" "D4W_ROW_FUSED=1 D4W_ROW_THREADS=192" "D4W_ROW_FUSED=1 D4W_ROW_THREADS=128"; do
  echo "== $cfg"; env $cfg timeout 120 python scripts/gpu_tune_fk.py --one 2>&1 | tail -1
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python scripts/gpu_pcie_probe.py 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['ms_per_step'])"
