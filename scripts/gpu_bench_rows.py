"""Device-time of the other SURVEY 8 rows at BASELINE config-3 scale (10 000 ch x 120 000 samp fp32):
CUDA-event times, algorithmic bytes per SURVEY 8(d), achieved GB/s.  One JSON line per operator."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import das4whales_b200 as dw
from das4whales_b200 import synth, rows
from das4whales_b200.fk import FkFilter
NX, NS = int(os.environ.get("NX", 10000)), int(os.environ.get("NS", 120000))
DX, FS = 2.0419046878814697, 200.0

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def report(name, ms, alg_bytes, note=""):
    print(json.dumps({"op": name, "ms": round(ms, 3), "channels_per_s": round(NX / (ms * 1e-3)), "algorithmic_GB": round(alg_bytes / 1e9, 2),
                      "achieved_GBs": round(alg_bytes / (ms * 1e-3) / 1e9, 1), "note": note}), flush=True)

x = synth.synth_strain(NX, NS, seed=1234)
S = NX * NS
time = np.arange(NS) / FS
hf = dw.detect.gen_template_fincall(time, FS, 17.8, 28.8, 0.68)
lf = dw.detect.gen_template_fincall(time, FS, 14.7, 21.8, 0.78)

m = dw.dsp.hybrid_ninf_filter_design((NX, NS), [0, NX, 1], DX, FS, 1350., 1450., 3300, 3450, 14., 30.)
flt = FkFilter(m); y = torch.empty_like(x)
report("fk_filter_sparsefilt[hybrid_ninf mask, %d/%d rows kept]" % (flt.rows_kept, NX // 2 + 1), timeit(lambda: flt(x, out=y)), 24 * S,
       "the reference scripts' mask: Butterworth tails are non-zero for every wavenumber, so nothing can be pruned exactly")
del flt, y; torch.cuda.empty_cache()
report("row_stats", timeit(lambda: rows.row_stats(x)), 4 * S)
report("compute_cross_correlograms[HF+LF, one pass]", timeit(lambda: rows.cross_correlogram(x, [hf, lf])), 12 * S, "includes row_stats pass")
report("compute_cross_correlogram[HF]", timeit(lambda: rows.cross_correlogram(x, [hf])), 8 * S, "includes row_stats pass")
report("envelope |hilbert|", timeit(lambda: rows.envelope(x)), 8 * S)
report("snr_tr_array(env=True)", timeit(lambda: rows.snr(x, env=True)), 8 * S, "includes row_stats pass")
report("snr_tr_array(env=False)", timeit(lambda: rows.snr(x, env=False)), 8 * S, "includes row_stats pass")
sos = dw.dsp.butterworth_filter([8, [14, 30], "bp"], FS)
report("bp_filt (order-8 Butterworth filtfilt)", timeit(lambda: rows.sosfiltfilt(sos, x, padlen=51), reps=1), 8 * S, "latency-bound recursion, fp64 state")
sub = x[:2000].contiguous()
report("stft_mag[160, hop 8, bins 2..32] on 2000 ch", timeit(lambda: rows.stft_mag(sub, 160, 8, 2, 32)) * NX / 2000, 4 * S, "scaled from 2000 channels")
kern = {'f0': 27., 'f1': 17., 'dur': 0.8, 'bdwidth': 4.}
import io, contextlib
def sc():
    with contextlib.redirect_stdout(io.StringIO()):
        return dw.detect.compute_cross_correlogram_spectrocorr(sub, FS, [14., 30.], kern, 0.8, 0.95)
report("compute_cross_correlogram_spectrocorr on 2000 ch", timeit(sc, reps=1) * NX / 2000, 4 * S, "scaled from 2000 channels")
# ---- round 2 additions ------------------------------------------------------------------------------------------------
corr = rows.cross_correlogram(x, [hf])[0]
env = rows.envelope(corr)
thr = 0.5 * float(env.max())
report("find_peaks flags (detection threshold)", timeit(lambda: rows.find_peaks_flags(env, thr)), 5 * S)
flags = rows.find_peaks_flags(env, thr)
report("compact_picks (device)", timeit(lambda: rows.compact_picks(flags)), S, "%d picks" % int(flags.sum()))
del corr, env, flags; torch.cuda.empty_cache()
raw = (x * 5.0e4).round().to(torch.int32)
report("raw2strain int32 -> fp32", timeit(lambda: rows.raw2strain(raw, 1e-9)), 8 * S)
del raw; torch.cuda.empty_cache()
report("get_fx[nfft 8192] ", timeit(lambda: dw.dsp.get_fx(x[:, :8192].contiguous(), 8192)), 8 * NX * 8192)
with contextlib.redirect_stdout(io.StringIO()):
    t_gab = timeit(lambda: dw.improcess.gabor_detect(x, FS, DX, [0, NX, 1], threshold=9100., threshold2=150.), reps=1)
report("improcess.gabor_detect (trace2image .. masked trace)", t_gab, 12 * S, "envelope/std + minmax + 1/10 binning + 2 x 101x101 filter2D pair + upsample-multiply")
img = dw.improcess.binning(dw.improcess.trace2image(x), 1 / 10, 1 / 10)
up, down = dw.improcess.gabor_filt_design(74.77)
report("filter2D 101x101 on the binned image", timeit(lambda: dw.improcess.filter2D(img, None, up + down)), 8 * img.numel(),
       "%d x %d image, %.2f GFMA" % (img.shape[0], img.shape[1], img.numel() * 10201 / 1e9))
# ---- the whole device-side pipeline of one file (no H2D) -------------------------------------------------------------------
from das4whales_b200 import pipeline
raw = (x * 5.0e4).round().to(torch.int32)
pipe = pipeline.MfDetectPipeline(NX, NS, [0, NX, 1], DX, FS, 1e-9)
pipe.process_device(raw); torch.cuda.synchronize()
import time as _t
t0 = _t.perf_counter()
for _ in range(3):
    res = pipe.process_device(raw)
torch.cuda.synchronize()
report("pipeline.process_device (raw2strain .. picks, data resident)", (_t.perf_counter() - t0) / 3 * 1e3, 4 * S,
       "%d + %d picks" % (res["picks_hf"][1].numel(), res["picks_lf"][1].numel()))
pipe2 = pipeline.MfDetectPipeline(NX, NS, [0, NX, 1], DX, FS, 1e-9, prune_eps=1e-5)
pipe2.process_device(raw); torch.cuda.synchronize()
t0 = _t.perf_counter()
for _ in range(3):
    res = pipe2.process_device(raw)
torch.cuda.synchronize()
report("pipeline.process_device, f-k mask pruned at eps = 1e-5", (_t.perf_counter() - t0) / 3 * 1e3, 4 * S)
