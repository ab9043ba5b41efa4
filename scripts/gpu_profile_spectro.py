"""Two passes of the spectrogram-correlation detector on 1000 x 120000 (first = warm-up) for an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import das4whales_b200 as dw
from das4whales_b200 import synth
xd = synth.synth_strain(1000, 120000, seed=3)
xd = xd.cuda() if not xd.is_cuda else xd
kern = {'f0': 27., 'f1': 17., 'dur': 0.8, 'bdwidth': 4.}
for _ in range(2):
    dw.detect.compute_cross_correlogram_spectrocorr(xd, 200., [14., 30.], kern, 0.8, 0.95)
torch.cuda.synchronize()
