"""GPU diagnostic: per-shape and per-pass check of the f-k pipeline against NumPy."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import das4whales_b200 as dw
from das4whales_b200 import _lib, fk
from oracle import dsp_oracle as O

DX, FS = 2.0419046878814697, 200.0
shapes = [(40, 240), (45, 175), (38, 120), (64, 400), (100, 1200), (300, 6000), (1024, 4096), (96, 18000)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
rng = np.random.default_rng(0)
for nx, ns in shapes:
    x = rng.standard_normal((nx, ns)).astype(np.float32)
    sel = [0, nx, 1]
    try:
        mask = dw.dsp.fk_filter_design((nx, ns), sel, DX, FS)
        plan = fk.get_plan(nx, ns, 0)
        flt = fk.FkFilter(mask)
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty_like(xd)
        L = _lib.lib()
        ws = flt._ws()
        nact = flt.rows_kept
        W = ws[: nact * ns * 8].view(torch.float32).view(nact, ns, 2)
        mref = O.fk_filter_design((nx, ns), sel, DX, FS)
        ref = O.fk_filter_filt(x.astype(np.float64), mref)
        msg = f"{nx}x{ns} t1={plan.t1} t2={plan.t2} tile={plan.tile} rows={nact}: "
        # pass 1 alone: W[slot][t] = FFT over channels, kept rows
        flt.run_pass(1, xd, yd)
        torch.cuda.synchronize()
        Wc = torch.view_as_complex(W.contiguous()).cpu().numpy()
        X1 = np.fft.fft(x.astype(np.float64), axis=0)
        msym = O.fold_mask(mref)
        act = [k for k in range(nx // 2 + 1) if np.any(msym[k] != 0)]
        e1 = np.max(np.abs(Wc - X1[act])) / np.max(np.abs(X1[act])) if act else 0
        msg += f"P1 err {e1:.2e} (nact host {len(act)}) "
        y = flt(xd).cpu().numpy()
        e = np.max(np.abs(y - ref)) / np.max(np.abs(ref))
        msg += f"full err {e:.2e}"
        # repeatability
        y2 = flt(xd).cpu().numpy()
        msg += f" repeat-diff {np.max(np.abs(y - y2)):.1e}"
        print(msg, flush=True)
    except Exception as ex:
        print(f"{nx}x{ns}: EXC {type(ex).__name__}: {ex}", flush=True)
