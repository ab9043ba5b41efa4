#!/usr/bin/env python
"""bench.py -- headline benchmark of das4whales_b200 (contract: see the task statement).

Metric (BASELINE.json): DAS channels/s through the f-k filter, plus achieved HBM GB/s vs the
measured roofline.  Workload at N=1: BASELINE.json configs[1] -- synthetic 10 000 ch x 120 000
samp fp32, f-k filter only (dsp.fk_filter_design fan mask), 1 x B200.  At N>1 every rank
filters its own 10 000 x 120 000 file (the reference path shards by file/channel block with no
exchange in this mode): weak scaling, no data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NX, NS = 10000, 120000
DX, FS = 2.0419046878814697, 200.0
FAN = (1400.0, 1450.0, 3400.0, 3500.0)
ALGO_BYTES_PER_SAMPLE = 24          # 3 HBM round trips x (read + write) x 4 B
METRIC = "DAS channels/sec through f-k filter"
CPU_SAMPLE_NX = 250                 # bounded CPU sample: 250 channels x the full 120 000 samples


# ncu --set full dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r01d_fk_pipe.txt; P2/P4 at their
# algorithmic 2.60 GB each)
NCU_TRAFFIC = {3: {"step": 7.34e9 + 2.60e9 + 3.27e9 + 2.60e9 + 8.04e9, "p5": 8.04e9,
                   "src": "profiles/r01d_fk_pipe.txt (P1 7.34 + P5 8.04 GB) + profiles/r01c_fk_tma_radix25.txt (P3 3.27 GB)"},
               1: {"step": 21.1e9, "p5": 6.44e9, "src": "profiles/r01c_fk_tma_radix25.txt (P1 6.18 + P3 3.27 + P5 6.44 GB)"}}
COL_KERNEL = {0: "k_col_inv_dual", 1: "k_col_inv_tma", 2: "k_colB_inv_fused + k_colA_inv", 3: "k_col2_pipe<inverse>"}


def dominant(pass_ms, peak, scheme):
    """SURVEY 8(d) K3 (inverse pass back to real samples): 4 B read + 4 B written per (channel, sample)."""
    ms = pass_ms[4]
    alg = 8 * NX * NS
    ach = alg / (ms * 1e-3) / 1e9
    return {"name": COL_KERNEL.get(scheme, "?") + " (P5, C2R over channels)", "ms": round(ms, 4), "algorithmic_bytes": alg,
            "achieved": round(ach, 1), "frac": round(ach / peak, 4), "traffic": NCU_TRAFFIC.get(scheme, {}).get("p5")}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def count_since(self, t_begin):
        return sum(1 for t, _ in self.rows if t >= t_begin)

    def stop(self, t_begin=0.0, window="timed region"):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t < t_begin:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def cpu_sample_inputs():
    import numpy as np
    rng = np.random.default_rng(1234)
    return rng.standard_normal((CPU_SAMPLE_NX, NS))


def time_cpu_reference(steps, warmup):
    """The reference's CPU arithmetic (oracle port of dsp.fk_filter_filt, NumPy pocketfft, one
    thread -- the reference itself is single-threaded) on a bounded sample of the workload."""
    import numpy as np
    from oracle import dsp_oracle as O           # allowed here: cpu_baseline / --impl reference legs only
    x = cpu_sample_inputs()
    mask = O.fk_filter_design(x.shape, [0, CPU_SAMPLE_NX, 1], DX, FS, *FAN)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.fk_filter_filt(x, mask)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    return CPU_SAMPLE_NX / t, t


def time_cpu_all_cores():
    """Same arithmetic as the reference's fk_filter_filt with the two FFTs handed to scipy.fft on every host core
    (NOT the reference's code path -- numpy.fft is single-threaded -- reported next to it for scale)."""
    import numpy as np
    import scipy.fft as sfft
    from oracle import dsp_oracle as O
    x = cpu_sample_inputs()
    mask = np.asarray(O.fk_filter_design(x.shape, [0, CPU_SAMPLE_NX, 1], DX, FS, *FAN))
    workers = os.cpu_count() or 1
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        spec = sfft.fftshift(sfft.fft2(x, workers=workers))
        y = sfft.ifft2(sfft.ifftshift(spec * mask), workers=workers).real
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    del y
    return {"value": CPU_SAMPLE_NX / best, "unit": "channels/s", "cores": workers,
            "note": "scipy.fft.fft2/ifft2(workers=all) on the same sample; not the reference's own path"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, t = time_cpu_reference(steps, warmup)
    sample = f"{CPU_SAMPLE_NX} ch x {NS} samp float64 per step (1/{NX // CPU_SAMPLE_NX} of the workload's channels, full time axis)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "channels/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic {NX} ch x {NS} samp, f-k filter only (fk_filter_design fan mask)",
                       "note": "reference CPU path (oracle port of dsp.fk_filter_filt: numpy.fft.fft2 -> mask -> ifft2, "
                               "complex128, single-threaded like the reference)"},
            "cpu_baseline": {"value": val, "unit": "channels/s", "cores": 1, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count(), "all_cores_variant": time_cpu_all_cores()},
            "e2e": {"value": val, "unit": "channels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mf", action="store_true", help="skip the f-k + matched-filter (BASELINE configs[2]) leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import das4whales_b200 as dw
    from das4whales_b200 import _lib, synth
    from das4whales_b200.fk import FkFilter

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    L = _lib.lib()
    mask = dw.dsp.fk_filter_design((NX, NS), [0, NX, 1], DX, FS, *FAN)
    flt = FkFilter(mask)
    x = synth.synth_strain(NX, NS, seed=1234 + rank)
    y = torch.empty_like(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # nvidia-smi needs ~1 s to come up: start it before the warm-up
    for _ in range(warmup):
        flt(x, out=y)
    # ---- per-pass device times (CUDA events on the launching stream) -------------------
    names = ["p1_col_fwd", "p2_row_split", "p3_row_mid", "p4_row_unsplit", "p5_col_inv"]
    pass_ms = [0.0] * 5
    reps = 5
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(reps)]
    torch.cuda.synchronize()
    for r in range(reps):
        ev[r][0].record()
        for i in range(5):
            flt.run_pass(i + 1, x, y)
            ev[r][i + 1].record()
    torch.cuda.synchronize()
    for r in range(reps):
        for i in range(5):
            pass_ms[i] += ev[r][i].elapsed_time(ev[r][i + 1]) / reps

    # ---- the timed region: exactly K steps ------------------------------------------------
    n0 = L.d4w_launch_count()
    barrier()
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        flt(x, out=y)
    e1.record()
    barrier()
    launches = L.d4w_launch_count() - n0
    clocks = None
    if rank == 0:
        # nvidia-smi reports every 100 ms; a short timed region (K steps of ~8 ms) may see fewer than three reports, so
        # the identical step keeps running, untimed, until three reports have been taken under the same load
        window = "timed region"
        t_cont = time.perf_counter()
        while sampler.proc and sampler.count_since(t_begin) < 3 and time.perf_counter() - t_cont < 3.0:
            for _ in range(8):
                flt(x, out=y)
            torch.cuda.synchronize()
            window = "timed region + untimed continuation of the same step (region shorter than 3 nvidia-smi periods)"
        clocks = sampler.stop(t_begin, window)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / steps
    value = NX * world / (ms_step * 1e-3)

    # ---- BASELINE configs[2]: f-k filter + fin-whale matched filter (HF + LF templates, one pass over the filtered data)
    mf = None
    if not args.no_mf:
        import numpy as np
        tgrid = np.arange(NS) / FS
        tpls = [dw.detect.gen_template_fincall(tgrid, FS, 17.8, 28.8, 0.68), dw.detect.gen_template_fincall(tgrid, FS, 14.7, 21.8, 0.78)]
        def fk_mf():
            flt(x, out=y)
            return dw.detect.compute_cross_correlograms(y, tpls)
        outs = fk_mf(); del outs
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mf_steps = min(steps, 5)
        m0.record()
        for _ in range(mf_steps):
            outs = fk_mf(); del outs
        m1.record()
        barrier()
        mf_ms = m0.elapsed_time(m1) / mf_steps
        if world > 1:
            t = torch.tensor([mf_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mf_ms = float(t.item())
        mf = {"workload": "BASELINE configs[2]: the same matrix through the f-k filter, then detect.compute_cross_correlograms "
                          "with the HF (17.8-28.8 Hz, 0.68 s) and LF (14.7-21.8 Hz, 0.78 s) fin-whale templates",
              "value": NX * world / (mf_ms * 1e-3), "unit": "channels/s", "ms_per_step": mf_ms, "steps": mf_steps,
              "algorithmic_bytes": (24 + 12) * NX * NS, "achieved_gbs": round((24 + 12) * NX * NS / (mf_ms * 1e-3) / 1e9, 1)}
        torch.cuda.empty_cache()

    # ---- end to end through the public API with HOST buffers ---------------------------------
    # A stream of files: every step copies that step's strain matrix from pinned host memory, filters it
    # through das4whales_b200.dsp.fk_filter_filt and copies the result back to pinned host memory.
    # Consecutive steps are software-pipelined over three CUDA streams (H2D of file i+1 and D2H of file
    # i-1 run under the filter of file i; PCIe is full duplex), double-buffered on the device.
    e2e = None
    if not args.no_e2e:
        e2e_steps = min(steps, 6)
        hx = hy = None
        try:                                   # 3 x 4.8 GB of pinned host memory per rank
            hx = torch.empty((NX, NS), dtype=torch.float32, pin_memory=True)
            hy = [torch.empty((NX, NS), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            ok = 1
        except Exception as exc:               # noqa: BLE001 -- report, never hang the other ranks
            ok, why = 0, repr(exc)[:200]
        if world > 1:
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
    if not args.no_e2e and not ok:
        e2e = {"unavailable": "pinned host buffers could not be allocated on every rank"}
        hx = hy = None
    elif not args.no_e2e:
        hx.copy_(x)
        torch.cuda.synchronize()
        del y
        xd = [x, torch.empty_like(x)]
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        s_cmp = torch.cuda.current_stream()

        def run_pipeline(n):
            ev_in = [None, None]       # H2D of buffer b finished
            ev_use = [None, None]      # filter finished reading buffer b
            ev_out = [None, None]      # D2H into host buffer b finished
            outs = [None, None]
            for i in range(n):
                bsel = i % 2
                with torch.cuda.stream(s_h2d):
                    if ev_use[bsel] is not None:
                        s_h2d.wait_event(ev_use[bsel])
                    xd[bsel].copy_(hx, non_blocking=True)
                    ev_in[bsel] = torch.cuda.Event(); ev_in[bsel].record(s_h2d)
                s_cmp.wait_event(ev_in[bsel])
                out = dw.dsp.fk_filter_filt(xd[bsel], mask)            # the public call (tensor in -> tensor out)
                ev_use[bsel] = torch.cuda.Event(); ev_use[bsel].record(s_cmp)
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_use[bsel])
                    if ev_out[bsel] is not None:
                        s_d2h.wait_event(ev_out[bsel])
                    hy[bsel].copy_(out, non_blocking=True)
                    out.record_stream(s_d2h)
                    ev_out[bsel] = torch.cuda.Event(); ev_out[bsel].record(s_d2h)
                outs[bsel] = out
            torch.cuda.synchronize()

        run_pipeline(2)                                              # warm-up (allocator, plans)
        barrier()
        t0 = time.perf_counter()
        run_pipeline(e2e_steps)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": NX * world * e2e_steps / dt, "unit": "channels/s", "h2d_bytes_per_step": NX * NS * 4,
               "d2h_bytes_per_step": NX * NS * 4, "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "api": "das4whales_b200.dsp.fk_filter_filt(cuda tensor, FkMask); per step: pinned-host H2D of the input matrix, "
                      "filter, D2H of the filtered matrix to pinned host memory; steps pipelined over 3 streams (wall clock "
                      "around all steps incl. the final synchronize)"}
        del hx, hy, xd

    if rank == 0:
        peak, peak_src = measured_peak()
        algo_bytes = ALGO_BYTES_PER_SAMPLE * NX * NS
        achieved = algo_bytes / (ms_step * 1e-3) / 1e9
        traffic = flt.traffic_bytes()
        scheme = flt.plan.col_scheme
        kernels = {n: {"ms": round(pass_ms[i], 4), "actual_bytes": traffic[n],
                       "actual_gbs": round(traffic[n] / (pass_ms[i] * 1e-3) / 1e9, 1) if pass_ms[i] > 0 else None}
                   for i, n in enumerate(names)}
        line = {"metric": METRIC, "value": value, "unit": "channels/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"synthetic {NX} ch x {NS} samp fp32, f-k filter only (fk_filter_design fan mask "
                                       f"{FAN}), one matrix per GPU",
                           "l2": "inputs (4.8 GB) exceed the 126 MB L2; no flush needed",
                           "rows_kept": flt.rows_kept, "rows_total": NX // 2 + 1,
                           "plan": {"t1": flt.plan.t1, "t2": flt.plan.t2, "col_tile_samples": flt.plan.tile,
                                    "col_scheme": scheme}},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": NCU_TRAFFIC.get(scheme, {}).get("step"),
                             "traffic_source": "ncu --set full dram__bytes_read+write per launch, "
                                               + NCU_TRAFFIC.get(scheme, {}).get("src", "n/a") + " + P2/P4 at their algorithmic 2.60 GB each",
                             "dominant_kernel": dominant(pass_ms, peak, scheme),
                             "peak_source": peak_src,
                             "scope": "whole f-k filter = 5 kernels per step; achieved = 24 B/(channel*sample) algorithmic bytes "
                                      "(SURVEY 8d) / step time; actual_bytes per kernel below are lower because wavenumber rows "
                                      "with an identically-zero folded mask are never stored",
                             "kernels": kernels},
                "clocks": clocks}
        if e2e:
            line["e2e"] = e2e
        if mf:
            line["fk_plus_matched_filter"] = mf
        if not args.no_cpu_baseline and world == 1:
            cv, ct = time_cpu_reference(steps=1, warmup=0)
            line["cpu_baseline"] = {"value": cv, "unit": "channels/s", "cores": 1, "kind": "port",
                                    "sample": f"{CPU_SAMPLE_NX} ch x {NS} samp float64, one fk_filter_filt call ({ct:.1f} s); "
                                              "oracle port of the reference's numpy.fft path, single-threaded like the reference",
                                    "host_cores": os.cpu_count(), "all_cores_variant": time_cpu_all_cores()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
