set -x
timeout 300 python -u -m pytest tests/test_fk_gpu.py -x -q --timeout 200 2>&1 | tail -5
for cfg in "" "D4W_COLB_THREADS=128" "D4W_COLB_RA=16" "D4W_COLB_RA=16 D4W_COLB_THREADS=128" "D4W_COLB_RA=16 D4W_COLB_THREADS=96" "D4W_COLB_RA=25 D4W_COLB_THREADS=128" "D4W_COLB_RA=25" "D4W_COLB_FUSED=0 D4W_COLB_PLAN=16,25 D4W_COLB_THREADS=128"; do
  echo "== $cfg"; env $cfg timeout 200 python scripts/gpu_tune_fk.py --one 2>&1 | tail -1
done
