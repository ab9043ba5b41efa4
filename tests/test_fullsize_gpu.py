"""BASELINE-size parity (configs[1] and configs[2]: 10 000 channels x 120 000 samples) -- DIRECT comparison of the CUDA f-k
filter with the float64 answer for the whole matrix, computed on the same GPU by oracle/torch_oracle.py (reference masks
and fft2 -> mask -> ifft2 restated in float64 torch; pinned against the NumPy oracle on the CPU by test_oracle_golden.py).
Contract: max-norm error <= 1e-4 and l2 error <= 1e-5 relative to the float64 result.

Config 3 adds the HF + LF fin-whale matched filter and the envelope on the filtered matrix; those are checked on 64 full
120 000-sample rows against the float64 NumPy/SciPy oracle (rows are independent in these operators)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DX, FS = 2.0419046878814697, 200.0
NX, NS = 10000, 120000
FAN = (1400.0, 1450.0, 3400.0, 3500.0)
HYB = (1350., 1450., 3300, 3450, 14., 30.)           # the masks the reference scripts use (main_mfdetect.py:46-47)


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, total = torch.cuda.mem_get_info()
    if total < 100 * 2 ** 30:
        pytest.skip("needs a GPU with >= 100 GB for the float64 whole-matrix oracle")
    import das4whales_b200 as dw
    from das4whales_b200 import _lib, synth
    _lib.lib()
    x = synth.synth_strain(NX, NS, seed=1234)
    return dw, torch, x


def _errors(torch, y32, ref64):
    d = y32.to(torch.float64) - ref64
    den_max = float(ref64.abs().max())
    den_l2 = float(torch.linalg.vector_norm(ref64))
    return float(d.abs().max()) / den_max, float(torch.linalg.vector_norm(d)) / den_l2


def _full_check(dw, torch, x, mask, mask64, tapering=False):
    from das4whales_b200.fk import FkFilter
    from oracle import torch_oracle as TO
    y = FkFilter(mask)(x, tapering=tapering)
    x64 = x.to(torch.float64)
    if tapering:
        import scipy.signal as sps
        x64 *= torch.from_numpy(sps.windows.tukey(NS, alpha=0.03)).to(x.device)[None, :]
    ref = TO.fk_filter_filt(x64, mask64)
    del x64, mask64
    e = _errors(torch, y, ref)
    del ref
    torch.cuda.empty_cache()
    return y, e


def test_config2_fan_mask_direct(env):
    dw, torch, x = env
    from oracle import torch_oracle as TO
    mask = dw.dsp.fk_filter_design((NX, NS), [0, NX, 1], DX, FS, *FAN)
    m64 = TO.fk_filter_design((NX, NS), [0, NX, 1], DX, FS, *FAN, device=x.device)
    _, (emax, el2) = _full_check(dw, torch, x, mask, m64)
    print(f"config 2 fan mask: max-norm {emax:.2e}, l2 {el2:.2e}")
    assert emax <= 1e-4 and el2 <= 1e-5, (emax, el2)


def test_config2_hybrid_ninf_mask_direct(env):
    dw, torch, x = env
    from oracle import torch_oracle as TO
    mask = dw.dsp.hybrid_ninf_filter_design((NX, NS), [0, NX, 1], DX, FS, *HYB)
    m64 = TO.hybrid_ninf_filter_design((NX, NS), [0, NX, 1], DX, FS, *HYB, device=x.device)
    _, (emax, el2) = _full_check(dw, torch, x, mask, m64, tapering=True)
    print(f"config 2 hybrid_ninf mask (tapered): max-norm {emax:.2e}, l2 {el2:.2e}")
    assert emax <= 1e-4 and el2 <= 1e-5, (emax, el2)


def test_config2_hybrid_ninf_eps_pruned_error_bound(env):
    """Opt-in support pruning by threshold (`eps`): rows of the folded mask that never exceed eps are dropped.  The
    error against the exact float64 answer must stay inside the 1e-4 contract for the documented eps = 1e-5."""
    dw, torch, x = env
    from oracle import torch_oracle as TO
    from das4whales_b200.fk import FkFilter
    mask = dw.dsp.hybrid_ninf_filter_design((NX, NS), [0, NX, 1], DX, FS, *HYB)
    flt = FkFilter(mask, eps=1e-5)
    assert flt.rows_kept < 0.3 * (NX // 2 + 1)
    y = flt(x)
    m64 = TO.hybrid_ninf_filter_design((NX, NS), [0, NX, 1], DX, FS, *HYB, device=x.device)
    ref = TO.fk_filter_filt(x.to(torch.float64), m64)
    del m64
    emax, el2 = _errors(torch, y, ref)
    print(f"config 2 hybrid_ninf eps=1e-5 ({flt.rows_kept} rows kept): max-norm {emax:.2e}, l2 {el2:.2e}")
    assert emax <= 1e-4, (emax, el2)


def test_config3_matched_filter_and_envelope_on_full_rows(env):
    dw, torch, x = env
    from oracle import detect_oracle as D
    from das4whales_b200.fk import FkFilter
    mask = dw.dsp.fk_filter_design((NX, NS), [0, NX, 1], DX, FS, *FAN)
    y = FkFilter(mask)(x)
    tgrid = np.arange(NS) / FS
    tpls = [dw.detect.gen_template_fincall(tgrid, FS, 17.8, 28.8, 0.68), dw.detect.gen_template_fincall(tgrid, FS, 14.7, 21.8, 0.78)]
    corr = dw.detect.compute_cross_correlograms(y, tpls)
    env_hf = dw.detect.envelope(corr[0])
    rows = np.random.default_rng(7).choice(NX, size=64, replace=False)
    rows.sort()
    ridx = torch.from_numpy(rows).to(y.device)
    yr = y[ridx].cpu().numpy().astype(np.float64)
    for t, c in zip(tpls, corr):
        ref = D.compute_cross_correlogram(yr, t)
        got = c[ridx].cpu().numpy().astype(np.float64)
        emax = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        el2 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert emax <= 1e-4 and el2 <= 1e-5, (emax, el2)
    # envelope of the GPU correlogram rows vs scipy.signal.hilbert on the same rows
    cr = corr[0][ridx].cpu().numpy().astype(np.float64)
    eref = D.envelope(cr)
    egot = env_hf[ridx].cpu().numpy().astype(np.float64)
    assert np.max(np.abs(egot - eref)) / np.max(eref) <= 1e-4
    # picks on those rows: indices equal to SciPy's on the same (GPU) envelope rows
    thr = 0.5 * float(eref.max())
    picks = dw.detect.pick_times_env(corr[0][ridx].contiguous(), thr)
    import scipy.signal as sps
    for i in range(len(rows)):
        ref_idx = sps.find_peaks(egot[i].astype(np.float32), prominence=thr)[0]
        assert np.array_equal(picks[i], ref_idx)
