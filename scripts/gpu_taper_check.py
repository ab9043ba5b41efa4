"""fk_filter_filt(tensor, tapering=True): result unchanged, caller's tensor tapered in place (edge samples only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import das4whales_b200 as dw
x = torch.randn(200, 2400, device="cuda")
mask = dw.dsp.fk_filter_design((200, 2400), [0, 200, 1], 2.0419046878814697, 200.0)
a = x.clone(); y = dw.dsp.fk_filter_filt(a, mask, tapering=True)
b = x.clone(); dw.dsp.taper_data(b)
y0 = dw.dsp.fk_filter_filt(b.clone(), mask, tapering=False)
print("side effect equal:", torch.equal(a, b), " result diff:", float((y - y0).abs().max() / y0.abs().max()))
