"""Synthetic DAS strain matrices: unit Gaussian noise plus fin-whale-like
hyperbolic chirps with hyperbolic move-out across channels.  Generated on the device with a
seeded torch.Generator (plumbing, not hot path)."""
import math

import numpy as np
import scipy.signal as sp

DX = 2.0419046878814697   # OOI RCA channel spacing [m]
FS = 200.0


def synth_strain(nx, ns, seed=1234, device=None, calls_per_minute=8, dx=DX, fs=FS, amp=3.0):
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    x = torch.randn((nx, ns), generator=g, device=dev, dtype=torch.float32)
    ncalls = int(round(calls_per_minute * ns / fs / 60.0))
    if ncalls <= 0 or amp == 0:
        return x
    L = int(0.68 * fs)
    t = np.arange(L) / fs
    c = torch.from_numpy((sp.chirp(t, f0=28.8, f1=17.8, t1=0.68, method="hyperbolic") * np.hanning(L)).astype(np.float32)).to(dev)
    rng = np.random.default_rng(seed)
    ch = torch.arange(nx, device=dev, dtype=torch.float32)
    ar = torch.arange(L, device=dev)
    for _ in range(ncalls):
        c0 = float(rng.integers(0, nx))
        t0 = float(rng.uniform(0, max(ns / fs - 10.0, 1.0)))
        delay = torch.sqrt(((ch - c0) * dx) ** 2 + 500.0 ** 2) / 1500.0
        i0 = ((t0 + delay) * fs).to(torch.int64)                       # [nx]
        ok = i0 + L <= ns
        rows = torch.nonzero(ok).flatten()
        if rows.numel() == 0:
            continue
        idx = i0[rows][:, None] + ar[None, :]
        x[rows[:, None].expand(-1, L), idx] += amp * c[None, :]
    return x


def plane_wave(nx, ns, k0, f0, device=None, phase=0.3):
    """cos(2*pi*(k0*c/nx + f0*t/ns) + phase): an eigenfunction of the f-k filter."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    c = torch.arange(nx, device=dev, dtype=torch.float64)
    t = torch.arange(ns, device=dev, dtype=torch.float64)
    a = torch.remainder(c * (k0 / nx), 1.0)[:, None] + torch.remainder(t * (f0 / ns), 1.0)[None, :]
    return torch.cos(2 * math.pi * a + phase).to(torch.float32)
