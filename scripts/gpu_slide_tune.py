"""Sliding STFT: runs per CTA (D4W_SLIDE_G) x sub-steps per run (D4W_SLIDE_Q) for a narrow and a wide band."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from das4whales_b200 import rows, synth
xd = synth.synth_strain(1000, 120000, seed=3)
xd = xd.cuda() if not xd.is_cuda else xd
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (b0, b1) in [(12, 24), (11, 41), (14, 20)]:
    for g in (8, 16):
        for q in (4, 6, 8, 12):
            os.environ["D4W_SLIDE_G"] = str(g); os.environ["D4W_SLIDE_Q"] = str(q)
            try:
                t = timed(lambda: rows.stft_mag(xd, 160, 8, b0, b1))
                print(f"bins {b0}..{b1} G={g} Q={q}: {t:.3f} ms", flush=True)
            except Exception as e:
                print(f"bins {b0}..{b1} G={g} Q={q}: {str(e)[:80]}")
