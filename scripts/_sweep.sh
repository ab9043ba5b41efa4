for cfg in "D4W_PIPE_PF=0" "D4W_PIPE_PF=1" "D4W_PIPE_PF=2" "D4W_PIPE_PF=4" "D4W_COL_PIPE=0 D4W_COL_CHUNK_MB=0" "D4W_PIPE_PF=0"; do
  echo "== $cfg"; env $cfg timeout 120 python scripts/gpu_tune_fk.py --one 2>&1 | tail -1
done
