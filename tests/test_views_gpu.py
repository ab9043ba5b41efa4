"""GPU parity tests (through the C ABI) for the functions added in round 2: spectral views (dsp.get_fx, dsp.instant_freq),
the alternate correlators (detect.xcorr, nxcorr2d), process_corr, unequal-length shift_xcorr, device pick compaction --
against tests/golden/views.npz (outputs of the unmodified reference) and the float64 oracle."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import dsp_oracle as O, detect_oracle as D

pytestmark = pytest.mark.gpu
TOL = 2e-5
FS = 200.0


@pytest.fixture(scope="module")
def dw():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import das4whales_b200 as dw
    from das4whales_b200 import _lib
    _lib.lib()
    return dw


def test_get_fx_golden(dw, golden):
    g = golden("views")
    for nfft in (512, 600, 1000):          # crop, exact, zero-pad (numpy.fft.fft(a, n) semantics)
        fx = dw.dsp.get_fx(g["fx_x"], nfft)
        assert fx.shape == (5, nfft) and fx.dtype == np.float64
        e = rel_err(fx, g[f"fx_{nfft}"])
        assert e[0] <= TOL and e[1] <= TOL, (nfft, e)


def test_get_fx_long_rows_vs_oracle(dw):
    import torch
    rng = np.random.default_rng(5)
    x = rng.standard_normal((37, 12000)).astype(np.float32)
    for nfft in (12000, 16384, 4096):
        fx = dw.dsp.get_fx(torch.from_numpy(x).cuda(), nfft)
        assert fx.is_cuda and fx.dtype == torch.float32
        e = rel_err(fx.cpu().numpy(), O.get_fx(x.astype(np.float64), nfft))
        assert e[0] <= TOL, (nfft, e)


def test_instant_freq_golden(dw, golden):
    g = golden("views")
    fi = dw.dsp.instant_freq(g["if_x"], FS)
    assert fi.shape == g["if_y"].shape
    # phase differences of an fp32 analytic signal: absolute tolerance in Hz relative to the sampling rate
    assert np.max(np.abs(fi - g["if_y"])) <= 2e-4 * FS
    assert np.max(np.abs(fi[50:-50] - g["if_y"][50:-50])) <= 2e-5 * FS


def test_shift_xcorr_unequal_lengths(dw, golden):
    g = golden("views")
    for tag, y in (("ab", g["sx_b"]), ("ac", g["sx_c"])):
        r = dw.detect.shift_xcorr(g["sx_a"], y)
        ref = g["sx_" + tag]
        assert r.shape == ref.shape
        assert rel_err(r, ref)[0] <= TOL, tag
        rn = dw.detect.shift_nxcorr(g["sx_a"], y)
        assert rel_err(rn, g["snx_" + tag])[0] <= TOL, tag


def test_shift_xcorr_long_dense_template(dw):
    """equal-length dense operands far beyond the correlator's 2 500-tap limit (pieces of 2 400 taps, summed shifted)"""
    rng = np.random.default_rng(12)
    a, b = rng.standard_normal(25000), rng.standard_normal(25000)
    r = dw.detect.shift_xcorr(a, b)
    assert rel_err(r, D.shift_xcorr(a, b))[0] <= TOL


def test_shift_xcorr_accepts_cuda_tensors(dw):
    import torch
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal(500), rng.standard_normal(500)
    r = dw.detect.shift_xcorr(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert r.is_cuda
    assert rel_err(r.cpu().numpy(), D.shift_xcorr(a, b))[0] <= TOL


def test_detect_xcorr_and_nxcorr2d_golden(dw, golden):
    g = golden("views")
    t_scale, cv = dw.detect.xcorr(g["xc_t"], g["xc_f"], g["xc_S"], g["xc_tvec"], g["xc_fvec"], g["xc_ker"])
    assert np.array_equal(t_scale, g["xc_tscale"])
    assert cv.shape == g["xc_val"].shape and cv[0] == 0 and cv[-1] == 0
    assert rel_err(cv, g["xc_val"])[0] <= 1e-4
    nf = len(g["xc_f"])
    nx2 = dw.detect.nxcorr2d(g["xc_S"][:nf], g["xc_ker"])
    assert rel_err(nx2, g["nxc2d"])[0] <= 1e-4


def test_process_corr_and_pick_times_par(dw, golden):
    g = golden("views")
    pk = dw.detect.process_corr(g["pc_x"], 0.05)
    # the envelope is fp32 on the GPU: picks whose prominence sits within fp32 noise of the threshold may differ
    ref = g["pc_idx"]
    assert len(set(pk.tolist()) ^ set(ref.tolist())) <= 1
    mf = golden("matched_filter")
    par = dw.detect.pick_times_par(mf["corr_hf"], 0.05)
    env = dw.detect.pick_times_env(mf["corr_hf"], 0.05)
    assert all(np.array_equal(a, b) for a, b in zip(par, env))


def test_compact_picks_matches_nonzero(dw):
    import torch
    from das4whales_b200 import rows
    gen = torch.Generator(device="cuda").manual_seed(3)
    for shape in ((7, 100), (33, 4099), (300, 12000)):
        flags = (torch.rand(shape, device="cuda", generator=gen) < 0.01).to(torch.uint8)
        flags[0] = 0                                    # an empty row
        off, idx = rows.compact_picks(flags)
        nz = torch.nonzero(flags)
        assert int(off[-1]) == nz.shape[0]
        assert torch.equal(idx.to(torch.int64), nz[:, 1])
        counts = torch.bincount(nz[:, 0], minlength=shape[0])
        assert torch.equal((off[1:] - off[:-1]).to(torch.int64), counts)
    off, idx = rows.compact_picks(torch.zeros((5, 64), dtype=torch.uint8, device="cuda"))
    assert int(off[-1]) == 0 and idx.numel() == 0


def test_free_plans_keeps_outstanding_filters_valid(dw):
    """a FkMask / FkFilter that survives fk.free_plans() must not touch a destroyed plan."""
    from das4whales_b200 import fk
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 400)).astype(np.float32)
    mask = dw.dsp.fk_filter_design((64, 400), [0, 64, 1], 2.0419046878814697, FS)
    y0 = dw.dsp.fk_filter_filt(x, mask)
    fk.free_plans()
    y1 = dw.dsp.fk_filter_filt(x, mask)                  # new plan, new device table
    assert np.array_equal(y0, y1)
    fk.free_plans()
    for _ in range(3):                                   # plans created in between may re-use freed addresses
        dw.dsp.fk_filter_filt(rng.standard_normal((32, 200)).astype(np.float32),
                              dw.dsp.fk_filter_design((32, 200), [0, 32, 1], 2.0419046878814697, FS))
    assert np.array_equal(dw.dsp.fk_filter_filt(x, mask), y0)


def test_dense_mask_cache_sees_in_place_changes(dw):
    """a caller-owned dense mask that is modified in place must not hit the stale device table."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((64, 400)).astype(np.float32)
    m = np.asarray(O.fk_filter_design((64, 400), [0, 64, 1], 2.0419046878814697, FS)).copy()
    y0 = dw.dsp.fk_filter_filt(x, m)
    m *= 0.5
    y1 = dw.dsp.fk_filter_filt(x, m)
    assert rel_err(y1, 0.5 * y0)[0] <= 1e-6


def test_fkmask_has_the_sparse_coo_surface_the_reference_prints(dw):
    """tools.disp_comprate (tools.py:239-257) reads mask.data.nbytes and mask.todense(); scripts call it on every hybrid mask."""
    m = dw.dsp.hybrid_ninf_filter_design((64, 400), [0, 64, 1], 2.0419046878814697, FS, 1350., 1450., 3300, 3450, 14., 30.)
    dense = m.todense()
    assert m.data.ndim == 1 and m.data.size == np.count_nonzero(dense) == m.nnz
    assert m.data.nbytes > 0 and dense.size * dense.itemsize >= m.data.nbytes
