"""CPU: host-side logic of the drop-in layer that needs no GPU -- design-time mask construction
(bit-identical to the reference's golden masks), templates, pick bookkeeping, filter design and
the API surface (names / defaults) the reference's callers rely on."""
import inspect

import numpy as np

import das4whales_b200 as dw

DX, FS = 2.0419046878814697, 200.0


def test_dense_design_functions_match_reference_golden(golden):
    g = golden("masks")
    sel = [0, 40, 1]
    assert np.array_equal(dw.dsp.hybrid_filter_design((40, 240), sel, DX, FS, 1400., 1450., 15., 25.).todense(), g["hyb_40x240"])
    assert np.array_equal(dw.dsp.hybrid_filter_design((38, 120), [0, 38, 1], DX, FS, 1400., 1450., 15., 25.).todense(), g["hyb_38x120"])
    assert np.array_equal(dw.dsp.hybrid_gs_filter_design((40, 240), sel, DX, FS, 1400., 1450., 15., 25.).todense(), g["gs_40x240"])
    assert np.array_equal(dw.dsp.hybrid_ninf_gs_filter_design((40, 240), sel, DX, FS, 1400., 1450., 3400, 3500, 15., 25.).todense(),
                          g["ninfgs_40x240"])


def test_lazy_masks_have_reference_shape_and_api():
    # the reference's own tests only assert shapes (tests/test_dsp.py:21-83)
    for fn in (dw.dsp.fk_filter_design, dw.dsp.hybrid_ninf_filter_design, dw.dsp.hybrid_filter_design,
               dw.dsp.hybrid_gs_filter_design, dw.dsp.hybrid_ninf_gs_filter_design):
        m = fn((10, 10), [0, 1, 2], 1, 100)     # the reference tests' arguments
        assert m.shape == (10, 10) and m.ndim == 2 and hasattr(m, "todense")


def test_signatures_match_reference_defaults():
    sig = inspect.signature(dw.dsp.fk_filter_design)
    assert list(sig.parameters) == ["trace_shape", "selected_channels", "dx", "fs", "cs_min", "cp_min", "cp_max", "cs_max"]
    assert [sig.parameters[k].default for k in ("cs_min", "cp_min", "cp_max", "cs_max")] == [1400, 1450, 3400, 3500]
    sig = inspect.signature(dw.dsp.hybrid_ninf_filter_design)
    assert [sig.parameters[k].default for k in ("cs_min", "cp_min", "cp_max", "cs_max", "fmin", "fmax")] == [1400., 1450., 3400, 3500, 15., 25.]
    assert list(inspect.signature(dw.dsp.fk_filter_filt).parameters) == ["trace", "fk_filter_matrix", "tapering"]
    assert list(inspect.signature(dw.dsp.fk_filter_sparsefilt).parameters) == ["trace", "fk_filter_matrix", "tapering"]
    assert list(inspect.signature(dw.dsp.bp_filt).parameters) == ["data", "fs", "fmin", "fmax"]
    assert list(inspect.signature(dw.dsp.get_spectrogram).parameters) == ["waveform", "fs", "nfft", "overlap_pct"]
    assert list(inspect.signature(dw.dsp.snr_tr_array).parameters) == ["trace", "env"]
    assert list(inspect.signature(dw.detect.compute_cross_correlogram).parameters) == ["data", "template"]
    assert list(inspect.signature(dw.detect.gen_template_fincall).parameters) == ["time", "fs", "fmin", "fmax", "duration", "window"]
    assert list(inspect.signature(dw.detect.compute_cross_correlogram_spectrocorr).parameters) == \
        ["data", "fs", "flims", "kernel", "win_size", "overlap_pct"]
    assert list(inspect.signature(dw.dsp.get_fx).parameters) == ["trace", "nfft"]
    assert list(inspect.signature(dw.dsp.instant_freq).parameters) == ["channel", "fs"]
    assert list(inspect.signature(dw.detect.xcorr).parameters) == ["t", "f", "Sxx", "tvec", "fvec", "BlueKernel"]
    assert list(inspect.signature(dw.detect.nxcorr2d).parameters) == ["spectro", "kernel"]
    assert list(inspect.signature(dw.detect.process_corr).parameters) == ["corr", "threshold"]
    assert list(inspect.signature(dw.detect.pick_times_par).parameters) == ["corr_m", "threshold"]
    assert list(inspect.signature(dw.improcess.trace2image).parameters) == ["trace"]
    assert list(inspect.signature(dw.improcess.gabor_filt_design).parameters) == ["theta_c0", "plot"]
    assert list(inspect.signature(dw.improcess.binning).parameters) == ["image", "ft", "fx"]
    assert list(inspect.signature(dw.improcess.angle_fromspeed).parameters) == ["c0", "fs", "dx", "selected_channels"]
    sig = inspect.signature(dw.improcess.apply_smooth_mask)
    assert list(sig.parameters) == ["array", "mask", "sigma"] and sig.parameters["sigma"].default == 1.5
    # north-star aliases
    assert dw.dsp.bandpass is dw.dsp.bp_filt and dw.dsp.compute_spectrogram is dw.dsp.get_spectrogram
    assert dw.detect.matched_filter is dw.detect.compute_cross_correlogram


def test_templates_and_pick_bookkeeping(golden):
    g = golden("matched_filter")
    time = np.arange(1600) / FS
    assert np.array_equal(dw.detect.gen_template_fincall(time, FS, 17.8, 28.8, 0.68), g["tpl_hf"])
    assert np.array_equal(dw.detect.gen_template_fincall(time, FS, 14.7, 21.8, 0.78), g["tpl_lf"])
    assert np.array_equal(dw.detect.gen_linear_chirp(15., 25., 1.0, FS), g["lin_chirp"])
    # reference tests/test_detect.py style length checks
    assert len(dw.detect.gen_hyperbolic_chirp(15., 25., 1.0, 200)) == 200
    picks = [np.array([1, 5]), np.array([], dtype=int), np.array([7])]
    tp = dw.detect.convert_pick_times(picks)
    assert np.array_equal(tp, np.array([[0, 0, 2], [1, 5, 7]]))
    sel = dw.detect.select_picked_times(tp, 0.02, 0.03, FS)
    assert np.array_equal(sel[0], [0]) and np.array_equal(sel[1], [5])


def test_butterworth_and_taper_kat(golden):
    sos = dw.dsp.butterworth_filter([5, [10, 30], "bp"], FS)
    assert np.array_equal(sos, golden("iir")["sos_bp5"])
    # reference KAT tests/test_dsp.py:85-88 (host path of taper_data)
    t = dw.dsp.taper_data(np.array([[1., 2, 3, 4, 5], [1, 2, 3, 4, 5]]))
    assert np.array_equal(t, np.array([[0., 2, 3, 4, 0], [0, 2, 3, 4, 0]]))
    x = np.random.default_rng(0).standard_normal((3, 400))
    y = x.copy()
    dw.dsp._taper_edges_inplace(y)
    import scipy.signal as sp
    assert np.array_equal(y, x * sp.windows.tukey(400, alpha=0.03)[None, :])
    # tensor inputs of fk_filter_filt(tapering=True) get the same side effect from the edge samples only
    import torch
    for nt in (5, 8, 1201):
        a = np.random.default_rng(nt).standard_normal((3, nt)).astype(np.float32)
        t = torch.from_numpy(a.copy())
        dw.dsp._taper_edges_inplace(t)
        full = torch.from_numpy(a.copy()) * torch.from_numpy(sp.windows.tukey(nt, alpha=0.03)).to(torch.float32)[None, :]
        assert torch.equal(t, full)


def test_gabor_kernel_design_is_the_opencv_kernel(golden):
    """improcess.gabor_filt_design restates cv2.getGaborKernel on the host (no OpenCV in the product path)."""
    g = golden("gabor")
    up, down = dw.improcess.gabor_filt_design(float(g["theta"]))
    assert up.shape == (101, 101) and np.max(np.abs(up - g["up"])) <= 1e-14 and np.array_equal(down, np.flipud(up))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs first) prints one JSON line with the agreed keys, and the
    product arm refuses to run without a GPU instead of falling back to the CPU."""
    import json, os, subprocess, sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "channels/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and "sample" in line["cpu_baseline"]
    assert line["cpu_baseline"]["single_thread"]["cores"] == 1 and line["cpu_baseline"]["single_thread"]["value"] > 0
    assert line["steps"] >= 1 and line["config"]["workload"].startswith("synthetic 10000 ch x 120000 samp")
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    import torch
    if not torch.cuda.is_available():
        r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                            capture_output=True, text=True, timeout=300)
        assert r2.returncode != 0 and "{\"metric\"" not in r2.stdout


def test_pick_list_helpers_match_oracle():
    """convert_pick_times / select_picked_times are host-side list handling: same result as the restated reference."""
    import das4whales_b200 as dw
    from oracle import detect_oracle as D
    picks = [np.array([3, 10, 11], dtype=np.int64), np.empty(0, dtype=np.int64), np.array([7], dtype=np.int64), np.empty(0, dtype=np.int64)]
    got, ref = dw.detect.convert_pick_times(picks), D.convert_pick_times(picks)
    assert got.shape == (2, 4) and np.array_equal(got, ref)
    assert dw.detect.convert_pick_times([np.empty(0, dtype=np.int64)] * 3).shape == (2, 0)
    sel = dw.detect.select_picked_times(got, 0.02, 0.055, 200.0)
    assert np.array_equal(sel[0], [0, 0, 2]) and np.array_equal(sel[1], [10, 11, 7])
