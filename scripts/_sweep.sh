timeout 900 python -u -m pytest tests/test_rows_gpu.py -x -q --timeout 300 2>&1 | tail -8
timeout 300 python - <<'PY'
import torch, time, numpy as np
import das4whales_b200 as dw
from das4whales_b200 import rows, synth
nx, ns = 10000, 120000
x = synth.synth_strain(nx, ns, seed=3)
env = rows.envelope(x)
for thr in (0.5, 3.0, 6.0):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fl = rows.find_peaks_flags(env, thr)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pk = rows.find_peaks(env, thr)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"find_peaks thr={thr}: kernels {1e3*(t1-t0):.2f} ms, with nonzero+D2H+split {1e3*(t2-t1):.1f} ms, picks {sum(len(p) for p in pk)}")
raw = torch.randint(-2**20, 2**20, (nx, ns), dtype=torch.int32, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); s = rows.raw2strain(raw, 1e-9); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"raw2strain int32->f32: {1e3*(t1-t0):.2f} ms ({nx*ns*12/(t1-t0)/1e9:.0f} GB/s of 12 B/sample)")
PY
