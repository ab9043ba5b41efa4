"""File-level matched-filter detection pipeline on one GPU (BASELINE config 5: one file per GPU).

The device-side equivalent of scripts/main_mfdetect.py:42-103 of the reference:

    raw counts (int32, as stored in the HDF5 file)           data_handle.load_das_data / raw2strain   (data_handle.py:157-214)
      -> strain (demean, scale)                               d4w_raw2strain
      -> band-pass 14-30 Hz, zero phase                       dsp.bp_filt                              (main_mfdetect.py:53)
      -> f-k filter, hybrid_ninf mask                         dsp.fk_filter_sparsefilt                 (:46-47, :55)
      -> HF + LF fin-whale matched filter (one pass)          detect.compute_cross_correlogram x 2     (:72-80)
      -> detection threshold 0.5 * max correlation            (:83, :95)
      -> Hilbert envelope + prominence peak picking           detect.pick_times_env x 2                (:98-99)
      -> (channel, sample) pick lists                         detect.convert_pick_times                (:102-103)

Only the raw counts go up (pinned int32, 4 bytes per sample) and only the picks come down; every intermediate matrix
stays in HBM.  `MfDetectPipeline.stream(files)` overlaps the H2D copy of file i+1 with the processing of file i.
"""
import numpy as np

from . import detect as _detect
from . import dsp as _dsp
from . import fk as _fk
from . import rows as _rows


class MfDetectPipeline:
    def __init__(self, nx, ns, selected_channels, dx, fs, scale_factor, device=None, fmin=14., fmax=30.,
                 mask_speeds=(1350., 1450., 3300, 3450), hf=(17.8, 28.8, 0.68), lf=(14.7, 21.8, 0.78), thres_frac=0.5,
                 prune_eps=0.0, bandpass=True):
        import scipy.signal as sp
        torch = _fk._torch()
        self.torch = torch
        self.nx, self.ns, self.fs = int(nx), int(ns), float(fs)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.scale_factor = float(scale_factor)
        self.thres_frac = float(thres_frac)
        self.bandpass = bool(bandpass)
        with torch.cuda.device(self.device):
            self.mask = _dsp.hybrid_ninf_filter_design((nx, ns), selected_channels, dx, fs, *mask_speeds, fmin=fmin, fmax=fmax)
            self.fk = _fk.FkFilter(self.mask, device=self.device, eps=prune_eps)
        self.sos = sp.butter(8, [fmin / (fs / 2), fmax / (fs / 2)], "bp", output="sos")          # dsp.bp_filt (dsp.py:878)
        time = np.arange(ns) / fs
        self.templates = [_detect.gen_template_fincall(time, fs, *hf), _detect.gen_template_fincall(time, fs, *lf)]
        self._raw = [None, None]
        self._h2d = None

    # ---- device side -------------------------------------------------------------------------------------------------
    def process_device(self, raw, with_snr=False):
        """raw: int32 (or float32) CUDA tensor [nx, ns] of interrogator counts.  Returns a dict of DEVICE results:
        picks_hf / picks_lf = (offsets int32 [nx + 1], idx int32 [n]) , maxv (0-dim tensor), optionally snr_hf / snr_lf."""
        torch = self.torch
        with torch.cuda.device(self.device):
            x = _rows.raw2strain(raw, self.scale_factor)
            if self.bandpass:
                x = _rows.sosfiltfilt(self.sos, x, padlen=3 * 17)
            y = self.fk(x, out=x)
            corr = _rows.cross_correlogram_chunked(y, self.templates, normalize=True)
            del x, y
            rmax = torch.stack([_rows.row_max(c) for c in corr])                     # [2, nx]
            maxv = _rows.row_max(rmax.reshape(1, -1))[0]
            thres = self.thres_frac * float(maxv.item())                             # main_mfdetect.py:95 (one scalar D2H)
            out = {"maxv": maxv, "threshold": thres}
            for name, c, thr in (("hf", corr[0], thres * 0.9), ("lf", corr[1], thres)):
                env = _rows.envelope(c)
                out["picks_" + name] = _rows.find_peaks_device(env, thr)
                del env
                if with_snr:
                    out["snr_" + name] = _rows.snr(c, env=True)
            return out

    @staticmethod
    def picks_to_host(picks):
        """(offsets, idx) on the device -> array([[channel ...], [sample ...]]) like detect.convert_pick_times."""
        offsets, idx = picks
        off = offsets.cpu().numpy().astype(np.int64)
        tt = idx.cpu().numpy().astype(np.int64)
        ch = np.repeat(np.arange(len(off) - 1, dtype=np.int64), np.diff(off))
        return np.asarray((ch, tt))

    # ---- host side ---------------------------------------------------------------------------------------------------
    def _staging(self, i, like):
        torch = self.torch
        if self._raw[i] is None or self._raw[i].dtype != like.dtype:
            self._raw[i] = torch.empty((self.nx, self.ns), dtype=like.dtype, device=f"cuda:{self.device}")
        return self._raw[i]

    def process_file(self, raw_host):
        """One file: raw_host = int32 / float32 ndarray or (pinned) CPU tensor [nx, ns].  Returns host pick arrays."""
        return next(self.stream([raw_host]))

    def stream(self, files):
        """Iterate over host arrays / tensors; yields {"picks_hf", "picks_lf", "maxv", "threshold"} per file with the H2D
        copy of the next file running on a second stream under the processing of the current one."""
        torch = self.torch
        if self._h2d is None:
            self._h2d = torch.cuda.Stream(device=self.device)
        it = iter(files)
        cur_stream = torch.cuda.current_stream(self.device)

        def upload(i, h):
            t = h if isinstance(h, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(h))
            if t.dtype not in (torch.int32, torch.float32):
                t = t.to(torch.float32)
            dst = self._staging(i, t)
            with torch.cuda.stream(self._h2d):
                self._h2d.wait_stream(cur_stream)                 # the buffer's previous consumer has finished
                dst.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._h2d)
            return dst, ev

        nxt = next(it, None)
        if nxt is None:
            return
        pending = upload(0, nxt)
        i = 0
        while pending is not None:
            dst, ev = pending
            nxt = next(it, None)
            # enqueue the next file's copy first: it waits (on the copy stream) for the kernels already queued -- the
            # previous user of that staging buffer -- and then runs under this file's processing
            pending = upload((i + 1) % 2, nxt) if nxt is not None else None
            cur_stream.wait_event(ev)
            res = self.process_device(dst)
            yield {"picks_hf": self.picks_to_host(res["picks_hf"]), "picks_lf": self.picks_to_host(res["picks_lf"]),
                   "maxv": float(res["maxv"].item()), "threshold": res["threshold"]}
            i += 1


def process_file(raw, metadata, selected_channels, **kw):
    """Convenience wrapper: raw [nx, ns] counts + the reference's metadata dict (data_handle.get_acquisition_parameters:
    fs, dx, scale_factor) -> picks of the HF and LF fin-whale notes."""
    nx, ns = raw.shape
    pipe = MfDetectPipeline(nx, ns, selected_channels, metadata["dx"], metadata["fs"], metadata["scale_factor"], **kw)
    return pipe.process_file(raw)
