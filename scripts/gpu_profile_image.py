"""One call of the image-domain detector and of the final matched filter so that `ncu -k regex:...` captures each kernel once:
    ncu --set full --clock-control none -k regex:"k_xcorr_pfa|k_filter2d|k_resize|k_mask_upsample|k_minmax|k_scale_pixels|k_peaks|k_row_fftmag" \
        -o gpurun_out/r02_image python scripts/gpu_profile_image.py"""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import das4whales_b200 as dw
from das4whales_b200 import rows, synth

NX, NS, FS, DX = int(os.environ.get("PROF_NX", 2000)), 120000, 200.0, 2.0419046878814697
x = synth.synth_strain(NX, NS, seed=5)
t = np.arange(NS) / FS
tpls = [dw.detect.gen_template_fincall(t, FS, 17.8, 28.8, 0.68), dw.detect.gen_template_fincall(t, FS, 14.7, 21.8, 0.78)]
hf, lf = dw.detect.compute_cross_correlograms(x, tpls)                    # k_row_stats, k_xcorr_pfa
env = rows.envelope(hf)
off, idx = rows.find_peaks_device(env, 0.5 * float(env.max()))            # k_peak_*, k_peaks_count / k_scan_offsets / k_peaks_fill
with contextlib.redirect_stdout(io.StringIO()):
    masked = dw.improcess.gabor_detect(x, FS, DX, [0, NX, 1], threshold=9100., threshold2=150.)   # k_minmax, k_scale_pixels, k_resize_aa_*, k_filter2d, k_mask_upsample_mul
fx = dw.dsp.get_fx(x[:, :8192].contiguous(), 8192)                        # k_row_fftmag
torch.cuda.synchronize()
print("ok")
