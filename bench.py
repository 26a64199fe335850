#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 gridding hot path (contract: see the task brief / DESIGN.md).

Workload (BASELINE.json configs[1]): make_imaging_weight Briggs (robust 0.5) + standard PS gridding of a
synthetic ALMA-like set -- 903 baselines x 500 integrations x 128 channels x 2 pol = 115.6 M samples per GPU,
4096^2 grid, support 7, oversampling 100, fp32 data/grid (fp64 index math), continuum (mfs) imaging.
One STEP = zero the accumulators, density grid (A2), Briggs factors (A3), weight degrid (A4), standard
gridding of vis * imaging weight (A1) -- and, for N > 1, the two NCCL reductions the path needs (all-reduce of
the density before the degrid, reduce of the uv-grid to rank 0).  Steps are issued through
cngi_prototype_b200.distributed.ContinuumPipeline, which overlaps the collectives of one step with the kernels of
its neighbours (double-buffered accumulators); the pipeline is drained inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line.  `value` is device-resident throughput (inputs in HBM), `e2e` the same step through the
public API from pinned HOST buffers with H2D/D2H copies inside the timed region, `roofline` the dominant kernel
(std_grid_window) against the measured HBM peak, `cpu_baseline` the oracle port of the reference on host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_UV = 4096
SUPPORT, OVERSAMPLING = 7, 100
IW_PARMS = dict(weighting="briggs", robust=0.5)
METRIC = "visibilities gridded/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-time", type=int, default=500)
    ap.add_argument("--n-chan", type=int, default=128)
    ap.add_argument("--n-uv", type=int, default=N_UV)
    ap.add_argument("--chan-mode", default="continuum", choices=["continuum", "cube"])
    ap.add_argument("--cpu-sample-times", type=int, default=0, help="integrations in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--side-stream", action="store_true",
                    help="issue the imaging-weight chain of step k+1 on a concurrent high-priority stream (measured: 2.50 "
                         "instead of 2.59 ms/step, but the gridding kernel's own time is then measured under contention)")
    return ap.parse_args()


def workload_name(a):
    return ("C2 ALMA-like: Briggs(0.5) imaging weights + standard PS gridding, 903 bl x %d t x %d ch x 2 pol, "
            "%d^2 grid, S=7, os=100, %s" % (a.n_time, a.n_chan, a.n_uv, a.chan_mode))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def red_peak(footprint_bytes):
    """Reduction ("atomic") ceiling of this GPU, measured live: G 32-byte sectors/s of REDG.E.ADD.F32x2 in the shape the
    gridding kernel flushes with (8 lanes on 64 contiguous bytes, scattered over a buffer as large as the uv-grid)."""
    import torch
    from cngi_prototype_b200 import _lib
    from cngi_prototype_b200._devutil import ptr, stream
    n_cells = int(footprint_bytes // 8)
    buf = torch.zeros(n_cells, dtype=torch.complex64, device="cuda")
    blocks, per_thread, best = 148 * 16, 256, float("inf")
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().cngi_b200_microbench_red(ptr(buf), n_cells, 1, blocks, per_thread, stream()), "microbench_red")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del buf
    return blocks * 8 * per_thread * 8 / (best * 1e-3) / 1e9


# ------------------------------------------------------------------------------------------------------
#  CPU arm: the oracle port of the reference's numba loops on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_step(O, d, gp, gp_iw, cgk, n_threads):
    """Same step as the GPU arm, reference semantics, fp64 (the reference always computes in fp64)."""
    rho, sw = O._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp_iw,
                                              n_threads=n_threads)
    bf = O._calculate_briggs_parms(rho, sw, IW_PARMS)
    iw = O._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho, (0, 1), (2, 3)), d["uvw"], d["weight"], bf,
                                                      d["freq_chan"], gp_iw, n_threads=n_threads)
    g, s = O._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw, d["freq_chan"], cgk, gp, n_threads=n_threads)
    return g, s


def cpu_arm(a, steps, warmup, sample_times=0, budget_s=150.0):
    """Times `steps` steps of the CPU path over the first `sample_times` integrations of the workload.
    sample_times = 0: sized from two short calibration steps (t = fixed + slope * integrations; the fixed part is the
    zero-fill and sum of the per-thread grids) so that steps + warmup fit in budget_s, capped at the whole workload."""
    from oracle import oracle as O
    from cngi_prototype_b200 import synth
    O.build()
    cores = os.cpu_count() or 1
    n_threads = max(1, min(cores, 32))   # continuum: one private 4096^2 c128 grid per thread (as the reference's chunks)
    full = synth.config_c2(n_time=a.n_time if not sample_times else sample_times, n_chan=a.n_chan, dtype="f64", shard=0)
    cgk = O._create_prolate_spheroidal_kernel_1D(OVERSAMPLING, SUPPORT)
    gp = synth.grid_parms_for(a.n_uv, full["cell"], chan_mode=a.chan_mode)
    gp_iw = synth.grid_parms_for(a.n_uv, full["cell"], chan_mode=a.chan_mode, support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)

    def first(n):
        d = dict(full)
        for k in ("vis", "uvw", "weight"):
            d[k] = full[k][:n]
        return d

    def run(d):
        t0 = time.perf_counter()
        cpu_step(O, d, gp, gp_iw, cgk, n_threads)
        return time.perf_counter() - t0

    n_all = full["weight"].shape[0]
    how = "fixed by --cpu-sample-times"
    if not sample_times:
        n_a, n_b = min(8, n_all), min(40, n_all)
        run(first(n_a))                                   # library load, thread start-up
        t_a, t_b = run(first(n_a)), run(first(n_b))
        slope = max((t_b - t_a) / max(n_b - n_a, 1), 1e-6)
        fixed = max(t_a - slope * n_a, 0.0)
        sample_times = int(max(n_b, min(n_all, (budget_s / max(steps + warmup, 1) - fixed) / slope)))
        how = "sized for %d+%d steps in %.0f s from calibration steps of %d and %d integrations (%.2f s fixed + %.1f ms per integration)" % (
            steps, warmup, budget_s, n_a, n_b, fixed, slope * 1e3)
    d = first(sample_times)
    n_samples = d["weight"].size
    for _ in range(warmup):
        run(d)
    dt = sum(run(d) for _ in range(steps)) / steps
    return dict(value=n_samples / dt, unit="vis/s", cores=n_threads, kind="port",
                sample="%d of %d integrations of the same workload (%.1f M samples/step; %s), C port of the reference "
                       "numba loops (oracle/cngi_oracle.c, fp64), %d pthreads over time chunks with private grids + "
                       "tree sum like the reference's dask graph; host has %d logical cores"
                       % (sample_times, a.n_time, n_samples / 1e6, how, n_threads, cores)), dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = cpu_arm(a, a.steps, min(a.warmup, 1), a.cpu_sample_times, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "vis/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": min(a.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "note": "CPU arm: each step is a bounded sample of the workload"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "vis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------------
#  GPU arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons / power DURING the timed region through NVML in a background thread
    (an `nvidia-smi -lms` loop in a subprocess perturbs the measured GPU: +30 % step time at 50 ms sampling)."""

    def __init__(self, gpu_index, period_s=0.1):
        self.rows, self.gpu, self.period, self._stop, self.thread, self.h = [], gpu_index, period_s, False, None, None

    def start(self):
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception as e:   # no NVML: fall back to one nvidia-smi query after the run
            self.h, self.err = None, str(e)

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip() != ""]
            if self.gpu < len(ids) and ids[self.gpu].strip().isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def _loop(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                reasons = get(self.h)
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.perf_counter(), sm, reasons, power))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self, t_begin, t_end):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: %s" % getattr(self, "err", "?")]}
        self._stop = True
        self.thread.join(timeout=2)
        nv = self.nv
        rows = [r for r in self.rows if t_begin <= r[0] <= t_end] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": float(self.max_sm), "reasons": ["no samples"]}
        sm = sorted(r[1] for r in rows)
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[3] for r in rows), "source": "NVML, %.0f ms period" % (self.period * 1e3)}


def run_b200(a):
    import torch
    import torch.distributed as dist
    from cngi_prototype_b200 import synth, _lib
    from cngi_prototype_b200._standard_grid import standard_grid
    from cngi_prototype_b200._imaging_weight import (imaging_weight_grid, calculate_briggs_parms,
                                                     _standard_imaging_weight_degrid_numpy_wrap)
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    d = synth.config_c2(n_time=a.n_time, n_chan=a.n_chan, dtype="f32", shard=rank)
    n_samples = d["weight"].size
    cgk = _create_prolate_spheroidal_kernel_1D(OVERSAMPLING, SUPPORT)
    gp = synth.grid_parms_for(a.n_uv, d["cell"], chan_mode=a.chan_mode)
    gp_iw = synth.grid_parms_for(a.n_uv, d["cell"], chan_mode=a.chan_mode, support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)
    n_ic = a.n_chan if a.chan_mode == "cube" else 1

    # pinned host copies (e2e source) and device-resident copies (value)
    H = {k: torch.as_tensor(d[k]).pin_memory() for k in ("vis", "uvw", "weight", "freq_chan")}
    T = {k: v.to(dev) for k, v in H.items()}
    cgk_t = torch.as_tensor(cgk).to(dev)
    # "visibility" = one (time, baseline, chan, pol) sample; SURVEY 8d asks for total and valid counts: valid = what the
    # gridder's mask keeps (finite non-zero vis * weight, finite uv; the synthetic uv all land inside the grid)
    ok_uv = torch.isfinite(T["uvw"][..., 0]) & torch.isfinite(T["uvw"][..., 1])
    wd = T["vis"] * T["weight"]
    n_valid = int((torch.isfinite(wd.real) & torch.isfinite(wd.imag) & (wd != 0) & ok_uv[:, :, None, None]).sum().item())
    del wd, ok_uv
    density = torch.empty((n_ic, 2, a.n_uv, a.n_uv), dtype=torch.float64, device=dev)
    dsw = torch.empty((n_ic, 2), dtype=torch.float64, device=dev)
    grid = torch.empty((n_ic, 2, a.n_uv, a.n_uv), dtype=torch.complex64, device=dev)
    gsw = torch.empty((n_ic, 2), dtype=torch.float64, device=dev)
    grid_host = torch.empty(grid.shape, dtype=grid.dtype).pin_memory()
    gsw_host = torch.empty(gsw.shape, dtype=gsw.dtype).pin_memory()
    grid_evs = []

    from types import SimpleNamespace
    from cngi_prototype_b200 import distributed as D
    ops = D.cuda_ops()

    def make_bufs():
        return SimpleNamespace(density=torch.empty((n_ic, 2, a.n_uv, a.n_uv), dtype=torch.float64, device=dev),
                               dsw=torch.empty((n_ic, 2), dtype=torch.float64, device=dev),
                               grid=torch.empty((n_ic, 2, a.n_uv, a.n_uv), dtype=torch.complex64, device=dev),
                               gsw=torch.empty((n_ic, 2), dtype=torch.float64, device=dev))

    # the sharding / collective control flow is the one tests/test_distributed_gloo.py exercises on CPU with gloo
    side = torch.cuda.Stream(device=dev, priority=-1) if a.side_stream else None
    pipe = D.ContinuumPipeline(ops, gp, gp_iw, IW_PARMS, cgk_t, make_bufs, side_stream=side)

    def grid_hook(what):   # CUDA events around the dominant kernel, on the stream it is launched on
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if what == "begin":
            grid_evs.append([ev, None])
        else:
            grid_evs[-1][1] = ev

    def step(src, record_kernel=False):
        pipe.step(src, grid_hook=grid_hook if record_kernel else None)

    copy_stream = torch.cuda.Stream(device=dev)
    n_chunks = 8
    bounds = [(a.n_time * i) // n_chunks for i in range(n_chunks + 1)]

    d2h_stream = torch.cuda.Stream(device=dev)
    e2e_bufs = [SimpleNamespace(grid=grid, gsw=gsw, grid_host=grid_host, gsw_host=gsw_host, d2h_done=None),
                SimpleNamespace(grid=torch.empty_like(grid), gsw=torch.empty_like(gsw), grid_host=torch.empty_like(grid_host).pin_memory(),
                                gsw_host=torch.empty_like(gsw_host).pin_memory(), d2h_done=None)]
    e2e_count = [0]

    def e2e_step():
        """Same step from pinned HOST buffers: chunked H2D on a copy stream overlapped with the kernels, D2H of the
        result on a third stream so that it overlaps the NEXT step's H2D (PCIe is full duplex); result buffers are
        double-buffered.  Device staging buffers (T) are reused; all copies are inside the timed region."""
        main = torch.cuda.current_stream()
        eb = e2e_bufs[e2e_count[0] & 1]
        e2e_count[0] += 1
        if eb.d2h_done is not None:
            main.wait_event(eb.d2h_done)   # the D2H that last read this grid buffer
        density.zero_(), dsw.zero_(), eb.grid.zero_(), eb.gsw.zero_()
        copy_stream.wait_stream(main)
        evs_w, evs_v = [], []
        with torch.cuda.stream(copy_stream):
            T["freq_chan"].copy_(H["freq_chan"], non_blocking=True)
            for i in range(n_chunks):   # weights + uvw first: the density pass needs only those
                sl = slice(bounds[i], bounds[i + 1])
                T["uvw"][sl].copy_(H["uvw"][sl], non_blocking=True)
                T["weight"][sl].copy_(H["weight"][sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs_w.append(ev)
            for i in range(n_chunks):
                sl = slice(bounds[i], bounds[i + 1])
                T["vis"][sl].copy_(H["vis"][sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                evs_v.append(ev)
        for i in range(n_chunks):
            sl = slice(bounds[i], bounds[i + 1])
            main.wait_event(evs_w[i])
            imaging_weight_grid(T["uvw"][sl], T["weight"][sl], T["freq_chan"], gp_iw, grid=density, sum_weight=dsw)
        if world > 1:
            dist.all_reduce(density)
            dist.all_reduce(dsw)
        bf = calculate_briggs_parms(density, dsw, IW_PARMS)
        iw = _standard_imaging_weight_degrid_numpy_wrap(density, T["uvw"], T["weight"], bf, T["freq_chan"], gp_iw,
                                                        kernel_side_layout=True)
        for i in range(n_chunks):
            sl = slice(bounds[i], bounds[i + 1])
            main.wait_event(evs_v[i])
            standard_grid(T["vis"][sl], T["uvw"][sl], iw[sl], T["freq_chan"], cgk_t, gp, False, True, grid=eb.grid,
                          sum_weight=eb.gsw)
        if world > 1:
            dist.reduce(torch.view_as_real(eb.grid), 0)
            dist.reduce(eb.gsw, 0)
        if rank == 0:
            d2h_stream.wait_stream(main)
            with torch.cuda.stream(d2h_stream):
                eb.grid_host.copy_(eb.grid, non_blocking=True)
                eb.gsw_host.copy_(eb.gsw, non_blocking=True)
                eb.d2h_done = torch.cuda.Event()
                eb.d2h_done.record(d2h_stream)

    def e2e_drain():
        torch.cuda.current_stream().wait_stream(d2h_stream)

    def timed(fn, steps, warmup, after=None, **kw):
        for _ in range(warmup):
            fn(**kw)
        if after:
            after()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn(**kw)
        if after:
            after()   # drain the software pipeline: the last step's stage B and collectives are inside the timed region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t_begin, t_end

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_step, t_begin, t_end = timed(lambda: step(T, record_kernel=True), a.steps, max(a.warmup, 3), after=pipe.flush)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    kern_ms = float(np.mean([e0.elapsed_time(e1) for (e0, e1) in grid_evs[-a.steps:] if e1 is not None]))

    e2e = None
    if not a.no_e2e:
        ms_e2e, _, _ = timed(e2e_step, a.steps, 3, after=e2e_drain)
        h2d = sum(H[k].numel() * H[k].element_size() for k in H)
        d2h = grid_host.numel() * grid_host.element_size() + gsw_host.numel() * 8
        e2e = {"value": world * n_samples / (ms_e2e * 1e-3), "unit": "vis/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "note": "pinned host buffers -> chunked H2D (copy stream) overlapped with kernels -> D2H of grid+sum_weight on a third stream, overlapping the next step's H2D; drained inside the timed region"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (std_grid_window), algorithmic bytes per SURVEY.md section 8d:
    # n_samples * (8 B vis + 4 B weight) + uvw + grid written once
    peak, peak_src = peaks()
    alg_bytes = n_samples * 12 + a.n_time * d["n_baseline"] * 24 + n_ic * 2 * a.n_uv * a.n_uv * 8
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic, red_sectors = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, red_sectors = tj.get("std_grid_dram_bytes_per_launch"), tj.get("std_grid_red_sectors_per_launch")
        except Exception:
            traffic = None
    # the second ceiling BASELINE.json's metric names: reductions into the grid (sectors per launch from the ncu
    # capture of this kernel -- a property of the input and the window, not of timing -- over the live kernel time)
    atomic = None
    if red_sectors and a.n_time == 500 and a.n_chan == 128 and a.chan_mode == "continuum":
        rp = red_peak(n_ic * 2 * a.n_uv * a.n_uv * 8)
        atomic = {"achieved": red_sectors / (kern_ms * 1e-3) / 1e9, "peak": rp, "unit": "Gsector/s",
                  "frac": red_sectors / (kern_ms * 1e-3) / 1e9 / rp, "sectors_per_launch": int(red_sectors),
                  "note": "REDG.E.ADD.F32x2 sectors per launch (ncu l1tex__t_sectors_pipe_lsu_mem_global_op_red, "
                          "profiles/r01_traffic.json) over the live kernel time; peak = cngi_b200_microbench_red, 8 lanes x "
                          "8 B contiguous, footprint = the uv-grid, measured in this run"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "std_grid_window_kernel<float,complex,S=7,PP=2>",
                "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src,
                "note": "not HBM bound: the issue slots (62 % active) and the shared-memory pipe limit it; FP32 floor 0.45 ms/launch (49 taps x 2 pol x 2 FMA per sample); see DESIGN.md section 4.1"}

    cb = None
    if not a.no_cpu_baseline and world == 1:   # reported on rank 0 at N = 1 only (the N > 1 runs are the scaling series)
        cb, _ = cpu_arm(a, 1, 1, a.cpu_sample_times, budget_s=12.0)   # the whole workload when one step fits in ~6 s

    line = {"metric": METRIC, "value": world * n_samples / (ms_step * 1e-3), "unit": "vis/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "samples_per_gpu": int(n_samples), "valid_samples_per_gpu": n_valid,
                       "l2": "inputs (1.4 GB/step) are larger than L2 (126 MB), no explicit flush",
                       "parallelism": ("time-sharded x%d, NCCL all-reduce(density plane 0) + reduce(grid), overlapped across steps" % world) if world > 1 else "single GPU"},
            "vis_tap_per_s": world * n_samples * SUPPORT * SUPPORT / (ms_step * 1e-3),
            "gridding_kernel_vis_per_s": n_samples / (kern_ms * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": 5 * a.steps,
            "gpu_launches_note": "per step: iw_grid, iw_sumsq, iw_briggs_finalize, iw_degrid, std_grid_window; the 128-thread "
                                 "uv_scale table kernel in front of A2 / A4 / A1 and the memsets are not counted",
            "roofline": roofline, "atomic_roofline": atomic,
            "cpu_baseline": cb}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    a = parse()
    if a.gpus > 1 and "RANK" not in os.environ:   # plain `python bench.py --gpus N`: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    # stdout carries exactly ONE JSON line (rank 0's): everything libraries print to fd 1 (NCCL's version banner, ...)
    # is sent to stderr, and the line is written to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
