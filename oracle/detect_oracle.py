"""oracle/detect_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Float64 NumPy/SciPy restatement of the DAS4Whales `detect` hot-path functions
(file:line relative to /root/reference/src/das4whales/detect.py).  Same rules as
oracle/dsp_oracle.py: imported only by tests/, smoke() and bench.py's CPU legs; pinned
against the unmodified reference by oracle/make_golden.py -> tests/golden/.
The spectrogram-correlation functions sit on librosa.stft, which is not installed here:
they use dsp_oracle.stft_librosa (restated, cross-checked against scipy.signal.stft).
"""
import numpy as np
import scipy.signal as sps

from .dsp_oracle import stft_librosa


def gen_linear_chirp(fmin, fmax, duration, sampling_rate):
    """detect.py:20-41 -- down-sweep fmax -> fmin."""
    t = np.arange(0, duration, 1 / sampling_rate)
    return sps.chirp(t, f0=fmax, f1=fmin, t1=duration, method="linear")


def gen_hyperbolic_chirp(fmin, fmax, duration, sampling_rate):
    """detect.py:44-65."""
    t = np.arange(0, duration, 1 / sampling_rate)
    return sps.chirp(t, f0=fmax, f1=fmin, t1=duration, method="hyperbolic")


def gen_template_fincall(time, fs, fmin=15., fmax=25., duration=1., window=True):
    """Hann-windowed hyperbolic chirp zero-padded to len(time) -- detect.py:68-93."""
    c = gen_hyperbolic_chirp(fmin, fmax, duration, fs)
    out = np.zeros(np.shape(time))
    out[: len(c)] = c * np.hanning(len(c)) if window else c
    return out


def shift_xcorr(x, y):
    """Positive-lag cross-correlation -- detect.py:96-112."""
    return sps.correlate(x, y, mode="full", method="fft")[len(x) - 1:]


def shift_nxcorr(x, y):
    """detect.py:115-137."""
    c = sps.correlate(x, y, mode="full", method="fft")
    return (c / (np.std(x) * np.std(y) * len(x)))[len(x) - 1:]


def compute_cross_correlogram(data, template):
    """Row-normalised matched filter -- detect.py:140-166.  Rows are demeaned and divided
    by the abs-max of the *raw* row (:157); the padded template is demeaned and divided by
    its raw abs-max (:158); positive lags of the full correlation per channel (:163-164)."""
    data = np.asarray(data, dtype=np.float64)
    nd = (data - data.mean(axis=1, keepdims=True)) / np.abs(data).max(axis=1, keepdims=True)
    tp = (template - np.mean(template)) / np.max(np.abs(template))
    out = np.empty_like(data)
    for i in range(data.shape[0]):
        out[i] = shift_xcorr(nd[i], tp)
    return out


def compute_cross_correlogram_direct(data, template):
    """Same quantity by the closed form of SURVEY.md App. A.3 (short FIR + suffix sum);
    an independent route used to cross-check the GPU algorithm's structure."""
    data = np.asarray(data, dtype=np.float64)
    ns = data.shape[1]
    nz = np.nonzero(template)[0]
    L = int(nz[-1]) + 1 if len(nz) else 0
    c = np.asarray(template[:L], dtype=np.float64)
    mu = c.sum() / ns
    m = np.max(np.abs(template))
    xt = (data - data.mean(axis=1, keepdims=True)) / np.abs(data).max(axis=1, keepdims=True)
    out = np.empty_like(xt)
    for i in range(xt.shape[0]):
        full = np.correlate(np.concatenate((xt[i], np.zeros(L - 1))), c, mode="valid")
        suffix = np.cumsum(xt[i][::-1])[::-1]
        out[i] = (full[:ns] - mu * suffix) / m
    return out


def envelope(corr_m):
    """|hilbert| per row, the quantity pick_times_env thresholds -- detect.py:192."""
    return np.abs(sps.hilbert(np.asarray(corr_m, dtype=np.float64), axis=1))


def pick_times_env(corr_m, threshold):
    """detect.py:169-195."""
    return [sps.find_peaks(np.abs(sps.hilbert(c)), prominence=threshold)[0] for c in corr_m]


def pick_times(corr_m, threshold):
    """detect.py:249-274."""
    return [sps.find_peaks(c, prominence=threshold)[0] for c in corr_m]


def convert_pick_times(peaks_indexes_m):
    """detect.py:277-303 -> array([[channel idx...],[time idx...]])."""
    ch = [i for i, p in enumerate(peaks_indexes_m) for _ in p]
    tt = [int(e) for p in peaks_indexes_m for e in p]
    return np.asarray((ch, tt))


# ------------------------------------------------------------- spectrogram correlation
def get_sliced_nspectrogram(trace, fs, fmin, fmax, nperseg, nhop):
    """|STFT| / max, rows with fmin <= f <= fmax -- detect.py:334-408."""
    s = np.abs(stft_librosa(trace, nperseg, nhop))
    nf, nt = s.shape
    tt = np.linspace(0, len(trace) / fs, num=nt)
    ff = np.linspace(0, fs / 2, num=nf)
    p = s / np.max(s)
    sel = np.where((ff >= fmin) & (ff <= fmax))
    return p[sel], ff[sel], tt


def buildkernel(f0, f1, bdwdth, dur, f, t, samp, fmin, fmax):
    """Hat-function hyperbolic-sweep kernel x Hann in time -- detect.py:411-492."""
    tvec = np.linspace(0, dur, np.size(np.nonzero((t < dur * 8) & (t > dur * 7))))
    x = f[:, None] - (f0 * f1 * dur / ((f0 - f1) * tvec[None, :] + f1 * dur))
    k = (1 - x ** 2 / bdwdth ** 2) * np.exp(-x ** 2 / (2 * bdwdth ** 2))
    return tvec, f, k * np.hanning(len(tvec))[None, :]


def xcorr2d(spectro, kernel):
    """detect.py:579-602."""
    c = sps.fftconvolve(spectro, np.flip(kernel, axis=1), mode="same", axes=1)
    m = np.sum(c, axis=0)
    m[m < 0] = 0
    return m / (np.median(spectro) * kernel.shape[1])


def spectrocorr_params(fs, flims, kernel, win_size, overlap_pct):
    """Parameter derivation of detect.py:680-696."""
    nperseg = int(win_size * fs)
    nhop = int(np.floor(nperseg * (1 - overlap_pct)))
    fmin, fmax = flims
    if fmax - kernel["f1"] < 2 * kernel["bdwidth"]:
        fmax = kernel["f1"] + 3 * kernel["bdwidth"]
    if kernel["f0"] - fmin < 2 * kernel["bdwidth"]:
        fmin = kernel["f0"] - 3 * kernel["bdwidth"]
    return nperseg, nhop, fmin, fmax


def compute_cross_correlogram_spectrocorr(data, fs, flims, kernel, win_size, overlap_pct):
    """detect.py:650-708 (raw `data` rows are used; norm_data at :678 is dead code)."""
    data = np.asarray(data, dtype=np.float64)
    nperseg, nhop, fmin, fmax = spectrocorr_params(fs, flims, kernel, win_size, overlap_pct)
    _, ff, tt = get_sliced_nspectrogram(data[0], fs, fmin, fmax, nperseg, nhop)
    _, _, ker = buildkernel(kernel["f0"], kernel["f1"], kernel["bdwidth"], kernel["dur"], ff, tt, fs, fmin, fmax)
    out = np.empty((data.shape[0], len(tt)))
    for i in range(data.shape[0]):
        s, _, _ = get_sliced_nspectrogram(data[i], fs, fmin, fmax, nperseg, nhop)
        out[i] = xcorr2d(s, ker)
    return out


def process_corr(corr, threshold):
    """detect.py:198-218"""
    return sps.find_peaks(np.abs(sps.hilbert(corr)), prominence=threshold)[0]


def nxcorr2d(spectro, kernel):
    """detect.py:544-576"""
    c = sps.correlate(spectro, kernel, mode="same", method="fft") / (np.std(spectro) * np.std(kernel) * spectro.shape[1])
    return np.max(c, axis=0)


def xcorr(t, f, Sxx, tvec, fvec, BlueKernel):
    """detect.py:605-647 -- explicit sliding dot product"""
    tsz, fsz = np.size(tvec), np.size(fvec)
    cv = np.zeros(np.size(t) - (tsz - 1))
    for i in range(np.size(t) - tsz + 1):
        cv[i] = np.sum(BlueKernel * Sxx[:fsz, i:i + tsz])
    cv /= (np.median(Sxx) * tsz)
    cv[0] = 0
    cv[-1] = 0
    cv[cv < 0] = 0
    return [t[int(tsz / 2) - 1:-int(np.ceil(tsz / 2))], cv]
