"""Per-channel (row) operators on the GPU: row statistics, overlap-save matched filter, Hilbert
envelope / SNR, forward-backward SOS IIR, batched STFT.  Thin host wrappers over libd4w.so."""
import numpy as np

from . import _lib


def _nyi(name):
    raise _lib.D4WError(f"das4whales_b200.rows.{name}: kernel not built yet")


def sosfiltfilt(sos, x, padlen=None):
    _nyi("sosfiltfilt")


def stft_mag(x, nfft, hop):
    _nyi("stft_mag")


def snr(x, env=False):
    _nyi("snr")


def envelope(x):
    _nyi("envelope")


def cross_correlogram(x, templates):
    _nyi("cross_correlogram")
