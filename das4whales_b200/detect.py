"""das4whales_b200.detect -- drop-in for the hot-path functions of `das4whales.detect`
(/root/reference/src/das4whales/detect.py): same names, positional order and defaults.
"""
import numpy as np
import scipy.signal as sp

from . import rows as _rows
from .dsp import _is_tensor, _to_device, _to_host64


# ---- templates (host side: tiny, keeps SciPy's chirp formula -- detect.py:20-93) --------
def gen_linear_chirp(fmin, fmax, duration, sampling_rate):
    """Linear down-sweep fmax -> fmin (reference: detect.py:20-41)."""
    t = np.arange(0, duration, 1 / sampling_rate)
    return sp.chirp(t, f0=fmax, f1=fmin, t1=duration, method="linear")


def gen_hyperbolic_chirp(fmin, fmax, duration, sampling_rate):
    """Hyperbolic down-sweep fmax -> fmin (reference: detect.py:44-65)."""
    t = np.arange(0, duration, 1 / sampling_rate)
    return sp.chirp(t, f0=fmax, f1=fmin, t1=duration, method="hyperbolic")


def gen_template_fincall(time, fs, fmin=15., fmax=25., duration=1., window=True):
    """Fin-whale call template, zero-padded to len(time) (reference: detect.py:68-93)."""
    chirp_signal = gen_hyperbolic_chirp(fmin, fmax, duration, fs)
    template = np.zeros(np.shape(time))
    n = len(chirp_signal)
    template[:n] = chirp_signal * np.hanning(n) if window else chirp_signal
    return template


def shift_xcorr(x, y):
    """Positive-lag cross-correlation of two 1-D arrays (reference: detect.py:96-112)."""
    x = np.asarray(x, dtype=np.float64)
    out = _rows.cross_correlogram(_to_device(x[None, :]), [np.asarray(y, dtype=np.float64)], normalize=False)[0]
    return _to_host64(out)[0]


def shift_nxcorr(x, y):
    """Std-normalised positive-lag cross-correlation (reference: detect.py:115-137)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return shift_xcorr(x, y) / (np.std(x) * np.std(y) * len(x))


def compute_cross_correlogram(data, template):
    """Matched filter of every channel against `template` (reference: detect.py:140-166):
    rows demeaned and divided by their raw abs-max, template (zero-padded to ns) demeaned and
    divided by its abs-max, positive lags of the full correlation."""
    xd = _to_device(data)
    out = _rows.cross_correlogram(xd, [np.asarray(template, dtype=np.float64)], normalize=True)[0]
    return out if _is_tensor(data) else _to_host64(out)


def compute_cross_correlograms(data, templates):
    """Several templates in ONE pass over the data (HF + LF notes of scripts/main_mfdetect.py:79-80)."""
    xd = _to_device(data)
    outs = _rows.cross_correlogram(xd, [np.asarray(t, dtype=np.float64) for t in templates], normalize=True)
    return outs if _is_tensor(data) else [_to_host64(o) for o in outs]


def envelope(corr_m):
    """|hilbert(row)| for every row -- the quantity pick_times_env thresholds (detect.py:192)."""
    y = _rows.envelope(_to_device(corr_m))
    return y if _is_tensor(corr_m) else _to_host64(y)


def pick_times_env(corr_m, threshold):
    """Peaks of the Hilbert envelope with prominence >= threshold (reference: detect.py:169-195).
    Envelope on the GPU; prominence search (branchy, ragged output) on the host -- SURVEY 8(f)."""
    env = envelope(corr_m)
    if _is_tensor(env):
        env = env.cpu().numpy()
    return [sp.find_peaks(e, prominence=threshold)[0] for e in env]


def pick_times(corr_m, threshold):
    """Peaks of the raw correlogram (reference: detect.py:249-274)."""
    c = corr_m.cpu().numpy() if _is_tensor(corr_m) else np.asarray(corr_m)
    return [sp.find_peaks(r, prominence=threshold)[0] for r in c]


pick_times_par = pick_times_env


def convert_pick_times(peaks_indexes_m):
    """list of per-channel index arrays -> array([[channel...],[time...]]) (detect.py:277-303)."""
    ch = [i for i, p in enumerate(peaks_indexes_m) for _ in p]
    tt = [e for p in peaks_indexes_m for e in p]
    return np.asarray((ch, tt))


def select_picked_times(idx_tp, tstart, tend, fs):
    """Keep picks with tstart <= t <= tend (reference: detect.py:306-330)."""
    keep = (idx_tp[1] >= tstart * fs) & (idx_tp[1] <= tend * fs)
    return (idx_tp[0][keep], idx_tp[1][keep])


# north-star aliases
matched_filter = compute_cross_correlogram
xcorr_templates = compute_cross_correlograms
