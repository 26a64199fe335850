#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 gridding hot path (contract: see the task brief / DESIGN.md section 5).

Workload (BASELINE.json configs[1]): make_imaging_weight Briggs (robust 0.5) + standard PS gridding of a
synthetic ALMA-like set -- 903 baselines x 500 integrations x 128 channels x 2 pol = 115.6 M samples per GPU,
4096^2 grid, support 7, oversampling 100, fp32 data/grid (fp64 index math), continuum (mfs) imaging.
One STEP = zero the accumulators, density grid (A2), Briggs factors (A3), weight degrid (A4), standard
gridding of vis * imaging weight (A1) -- and, for N > 1, the two NCCL reductions the path needs (all-reduce of
the density before the degrid, reduce of the uv-grid to rank 0).  Steps are issued through
cngi_prototype_b200.distributed.ContinuumPipeline, which overlaps the collectives of one step with the kernels of
its neighbours (double-buffered accumulators); the pipeline is drained inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line.  Keys beyond the base contract:
  value        device-resident throughput of the step (inputs in HBM)
  e2e          the same step through the PUBLIC API on HOST arrays -- imaging.make_imaging_weight + imaging.make_grid on
               numpy views of page-locked buffers (distributed.* for N > 1: the NCCL reductions are inside the calls),
               every H2D / D2H copy inside the timed region; `pcie_frac` = achieved H2D rate / the rate of a plain
               concurrent H2D probe at the same N
  parity       GPU grid / sum_weight of the workload against the multi-threaded CPU oracle on the SAME inputs
  roofline     the dominant kernel (std_grid_window) against the measured HBM peak; atomic_roofline: its reductions
  sustained    the same step looped for >= 3 s, with its own clock record
  value_f64    the same step at the reference's precision (complex128 / float64)
  configs      the other BASELINE configs' rows on this GPU: C1 (fp64 image / psf gridding), C3 (aperture gridders), C4 (degrid)
  cube         BASELINE config 5's per-GPU share (8192^2 padded to 9830^2, 128 channels per GPU, channel-sharded):
               time_split = 1 (no exchange) and, for N >= 2, time_split = 2 (NCCL sub-group grid reduce)
  cpu_baseline the reference's own numba loops (oracle/_ref, kind "reference") when staged, else the C port
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
PARITY_BARS = {"rel_err": 1e-5, "sum_weight_rel_err": 1e-6, "mask_equal": True}


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
    ap.add_argument("--cpu-engine", default="auto", choices=["auto", "reference", "port"],
                    help="CPU arm: the reference's numba loops (oracle/_ref) or the C port (oracle/cngi_oracle.c)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cube", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 row timings (N = 1 only)")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / value_f64 / pcie probe")
    ap.add_argument("--cube-chan-per-gpu", type=int, default=128)
    ap.add_argument("--cube-samples-per-gpu", type=float, default=2.5e9)
    ap.add_argument("--cube-chan-chunk", type=int, default=8)
    ap.add_argument("--cube-steps", type=int, default=2)
    ap.add_argument("--cube-overlap", type=int, default=0, choices=[0, 1],
                    help="time_split = 1: grid chunk j + 1 while chunk j is transformed on a side stream (two grid buffers)")
    ap.add_argument("--side-stream", action="store_true",
                    help="issue the imaging-weight chain of step k+1 on a concurrent high-priority stream")
    ap.add_argument("--fuse-weights", action="store_true",
                    help="form the imaging weights inside the gridder (cngi_b200_standard_grid_weighted) instead of the A4 pass")
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


def traffic_profile():
    """ncu-derived per-launch constants of the dominant kernel (a property of the input and the kernel, not of timing):
    the newest profiles/rNN_traffic.json."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                with open(path) as f:
                    return json.load(f), name
            except Exception:
                pass
    return {}, None


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
#  CPU arm: the reference's numba loops (oracle/_ref) or the C port of them (oracle/cngi_oracle.c) on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_step(O, d, gp, gp_iw, cgk, n_threads):
    """Same step as the GPU arm, reference semantics, fp64 (the reference always computes in fp64) -- C port."""
    rho, sw = O._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp_iw,
                                              n_threads=n_threads)
    bf = O._calculate_briggs_parms(rho, sw, IW_PARMS)
    iw = O._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho, (0, 1), (2, 3)), d["uvw"], d["weight"], bf,
                                                      d["freq_chan"], gp_iw, n_threads=n_threads)
    g, s = O._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw, d["freq_chan"], cgk, gp, n_threads=n_threads)
    return g, s


def pick_engine(want):
    """'reference' when the reference's numba loops are staged (oracle/_ref, or /root/reference in the build container)
    and numba imports; else the C port."""
    if want in ("auto", "reference"):
        try:
            from oracle import ref_numba
            if ref_numba.available():
                return "reference"
        except Exception:
            pass
        if want == "reference":
            raise SystemExit("--cpu-engine reference: oracle/_ref is not staged (run oracle/make_ref.sh) or numba is missing")
    return "port"


def cpu_arm(a, steps, warmup, sample_times=0, budget_s=150.0, engine="port", data=None, keep_result=False):
    """Times `steps` steps of the CPU path over the first `sample_times` integrations of the workload.
    sample_times = 0: sized from two short calibration steps (t = fixed + slope * integrations; the fixed part is the
    zero-fill and sum of the per-thread grids) so that steps + warmup fit in budget_s, capped at the whole workload.
    data: the arrays to use (fp64; default: a freshly generated fp64 set of the same geometry).
    Returns (cpu_baseline record, seconds per step, (grid, sum_weight, n_integrations) of the last step or None)."""
    from oracle import oracle as O
    from cngi_prototype_b200 import synth
    O.build()
    cores = os.cpu_count() or 1
    n_threads = max(1, min(cores, 32))   # continuum: one private 4096^2 c128 grid per thread (as the reference's chunks)
    full = data if data is not None else synth.config_c2(n_time=a.n_time if not sample_times else sample_times,
                                                         n_chan=a.n_chan, dtype="f64", shard=0)
    gp = synth.grid_parms_for(a.n_uv, full["cell"], chan_mode=a.chan_mode)
    gp_iw = synth.grid_parms_for(a.n_uv, full["cell"], chan_mode=a.chan_mode, support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)
    if engine == "reference":
        from oracle import ref_numba
        R = ref_numba.ReferenceStep(n_threads)
        cgk = R.cgk_1D(OVERSAMPLING, SUPPORT)
        step_fn = lambda d: R.step(d, gp, gp_iw, IW_PARMS, cgk)   # noqa: E731
        what = ("the reference's own numba loops (_standard_grid.py:242,466, unmodified, staged by oracle/make_ref.sh, "
                "numba %s), one call per time chunk from a %d-thread pool (nogil, as dask's threaded scheduler runs them) "
                "with private grids + pairwise tree sum (_tree_sum_list)" % (__import__("numba").__version__, n_threads))
    else:
        cgk = O._create_prolate_spheroidal_kernel_1D(OVERSAMPLING, SUPPORT)
        step_fn = lambda d: cpu_step(O, d, gp, gp_iw, cgk, n_threads)   # noqa: E731
        what = ("C port of the reference numba loops (oracle/cngi_oracle.c, fp64), %d pthreads over time chunks with "
                "private grids + tree sum like the reference's dask graph" % n_threads)

    def first(n):
        d = dict(full)
        for k in ("vis", "uvw", "weight"):
            d[k] = full[k][:n]
        return d

    last = [None]

    def run(d):
        t0 = time.perf_counter()
        last[0] = step_fn(d)
        dt = time.perf_counter() - t0
        if not keep_result:
            last[0] = None
        return dt

    n_all = full["weight"].shape[0]
    how = "fixed by --cpu-sample-times"
    if not sample_times:
        n_a, n_b = min(8, n_all), min(40, n_all)
        run(first(n_a))                                   # library load / numba JIT, thread start-up
        t_a, t_b = run(first(n_a)), run(first(n_b))
        slope = max((t_b - t_a) / max(n_b - n_a, 1), 1e-6)
        fixed = max(t_a - slope * n_a, 0.0)
        sample_times = int(max(n_b, min(n_all, (budget_s / max(steps + warmup, 1) - fixed) / slope)))
        how = "sized for %d+%d steps in %.0f s from calibration steps of %d and %d integrations (%.2f s fixed + %.1f ms per integration)" % (
            steps, warmup, budget_s, n_a, n_b, fixed, slope * 1e3)
    sample_times = min(sample_times, n_all)
    d = first(sample_times)
    n_samples = d["weight"].size
    for _ in range(warmup):
        run(d)
    dt = sum(run(d) for _ in range(steps)) / steps
    rec = dict(value=n_samples / dt, unit="vis/s", cores=n_threads, kind=engine,
               sample="%d of %d integrations of the same workload (%.1f M samples/step; %s), %s; host has %d logical cores"
                      % (sample_times, a.n_time, n_samples / 1e6, how, what, cores))
    res = None
    if keep_result and last[0] is not None:
        res = (last[0][0], last[0][1], sample_times)
    if engine == "reference":
        R.close()
    return rec, dt, res


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    engine = pick_engine(a.cpu_engine)
    cb, dt, _ = cpu_arm(a, a.steps, min(a.warmup, 1), a.cpu_sample_times, budget_s=150.0, engine=engine)
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
        return self

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

    def window(self, t_begin, t_end):
        """Clock record of [t_begin, t_end] (the sampler keeps running: several timed regions share one thread)."""
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: %s" % getattr(self, "err", "?")]}
        nv = self.nv
        rows = [r for r in list(self.rows) if t_begin <= r[0] <= t_end] or list(self.rows)[-1:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": float(self.max_sm), "reasons": ["no samples"]}
        sm = sorted(r[1] for r in rows)
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_min_mhz": float(sm[0]), "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[3] for r in rows), "source": "NVML, %.0f ms period" % (self.period * 1e3)}

    def stop(self):
        self._stop = True
        if self.thread is not None:
            self.thread.join(timeout=2)


def pcie_probe(dev, world, dist, n_bytes=1 << 30, reps=4):
    """Plain H2D and D2H copies of a page-locked buffer, all ranks at once (barrier before): the PCIe / host-memory
    ceiling the e2e arm sits under at this N.  Returns GB/s per rank (min over ranks) for each direction."""
    import torch
    host = torch.empty(n_bytes, dtype=torch.uint8).pin_memory()
    devb = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, dst, src in (("h2d", devb, host), ("d2h", host, devb)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = torch.tensor([n_bytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(gbs, op=dist.ReduceOp.MIN)
        out[name + "_gbs_per_gpu"] = float(gbs.item())
    del host, devb
    return out


def cube_record(a, dev, rank, world, dist, time_split, groups):
    """BASELINE config 5's per-GPU share through distributed.cube_imaging: image 8192^2, padded grid 9830^2 (= int(1.2 *
    8192) = 2 * 5 * 983: cuFFT takes its Bluestein path), `cube_chan_per_gpu` channels and `cube_samples_per_gpu` samples per
    GPU (1024 channels / 2e10 samples on 8 GPUs), fp32, channel-sharded (synthesis_imaging_cube.py:105-124).
    time_split = 1: a rank owns its channels end to end, no exchange.  time_split = 2: pairs of ranks share a channel
    block of twice the size, each grids half of the integrations, and every chunk of planes is summed onto the pair's
    root with an NCCL reduce over the sub-group before the FFT (the "NCCL grid reduce" of config 5); the root of a chunk
    rotates over the pair (distributed.cube_imaging(rotate_roots=True)) so that both GPUs run FFTs."""
    import torch
    from cngi_prototype_b200 import synth, distributed as D
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    n_img, n_pad = 8192, int(1.2 * 8192)
    cg, n_cg, tp, root = D.cube_layout(rank, world, time_split)
    n_chan_total = a.cube_chan_per_gpu * world
    n_chan_grp = a.cube_chan_per_gpu * time_split
    rng = np.random.default_rng(4321)
    bl = synth.baseline_vectors(synth.antenna_layout(43, 300.0, rng))
    n_bl = len(bl)
    n_time = max(8, int(round(a.cube_samples_per_gpu / (n_bl * n_chan_grp * 2))))       # integrations of THIS rank
    t0 = tp * n_time - (time_split * n_time) // 2
    ha = synth.EARTH_RATE * 6.0 * (t0 + np.arange(n_time))
    uvw = synth.uvw_tracks(bl, ha, np.deg2rad(-23.0))
    freq_all = np.linspace(345.0e9, 347.0e9, n_chan_total)
    freq = freq_all[cg * n_chan_grp:(cg + 1) * n_chan_grp]
    cell = 1.0 / (2.0 * np.max(np.linalg.norm(bl, axis=1)) * 347.0e9 / synth.C_LIGHT * 1.15)
    gp = synth.grid_parms_for(n_pad, cell, chan_mode="cube")
    gp["image_size"] = np.array([n_img, n_img], dtype=np.int64)
    cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(OVERSAMPLING, SUPPORT)).to(dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    shape = (n_time, n_bl, n_chan_grp, 2)
    vis = torch.empty(shape, dtype=torch.complex64, device=dev)
    wgt = torch.empty(shape, dtype=torch.float32, device=dev)
    tb = max(1, n_time // 16)
    for t in range(0, n_time, tb):   # generated on the device in time blocks (bounded transients)
        v = torch.view_as_real(vis[t:t + tb])
        v.normal_(generator=gen)
        wgt[t:t + tb].uniform_(0.5, 1.5, generator=gen)
        flagged = torch.rand(v.shape[:-1], device=dev, generator=gen) < 0.02
        v[flagged] = float("nan")
        del flagged
    d = {"vis": vis, "weight": wgt, "uvw": torch.as_tensor(uvw).to(dev), "freq_chan": torch.as_tensor(freq).to(dev)}
    n_samples = int(np.prod(shape))
    ops = D.cuda_ops()
    timer = D._PhaseTimer()

    def one(tm=None):
        return D.cube_imaging(ops, d, gp, cgk, chan_chunk=a.cube_chan_chunk, time_split=time_split, groups=groups,
                              presharded=True, timer=tm, keep_image=False, rotate_roots=True,
                              overlap=bool(a.cube_overlap) and time_split == 1)

    one()                                    # warm-up: cuFFT plan, allocator pools
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sw, owned = None, None
    for _ in range(a.cube_steps):
        out = one(timer)
        if out[1] is not None:
            sw, owned = out[1], out[-1]
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.cube_steps], device=dev, dtype=torch.float64)
    phases = {k: v / a.cube_steps for k, v in timer.ms().items()}
    ph = torch.tensor([phases.get("grid", 0.0), phases.get("reduce", 0.0), phases.get("image", 0.0)], device=dev,
                      dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    ok = None
    if sw is not None:   # every plane this rank transformed has a finite, positive sum of weights
        rows = torch.cat([sw[c0:c1] for c0, c1 in owned]) if isinstance(owned, list) else sw
        ok = bool(torch.isfinite(rows).all() and (rows > 0).all() and rows.numel() > 0)
    del d, vis, wgt
    torch.cuda.empty_cache()
    ms = float(ms.item())
    g_ms, r_ms, i_ms = (float(x) for x in ph.tolist())
    tj, tname = traffic_profile()
    rec = {"time_split": time_split, "ms_per_step": ms, "value": world * n_samples / (ms * 1e-3), "unit": "vis/s",
           "samples_per_gpu": n_samples, "n_chan_total": n_chan_total, "chan_per_group": n_chan_grp,
           "integrations_per_gpu": n_time, "chan_chunk": a.cube_chan_chunk, "steps": a.cube_steps, "warmup": 1,
           "phase_ms_max_over_ranks": {"grid": g_ms, "reduce_incl_wait_for_partner": r_ms, "image_fft_crop_correct": i_ms},
           "fft_share": i_ms / ms if ms else None,
           "gridding_vis_per_s_per_gpu": n_samples / (g_ms * 1e-3) if g_ms else None,
           "sum_weight_finite_positive": ok, "overlap": bool(a.cube_overlap) and time_split == 1}
    spp = tj.get("cube_red_sectors_per_sample")
    if spp and g_ms:
        rec["red_sectors_per_s"] = spp * n_samples / (g_ms * 1e-3)
        rec["red_note"] = "REDG sectors per sample from ncu (profiles/%s) x samples / grid phase; ceiling 40-65 G sectors/s (DESIGN 4.1)" % tname
    return rec


def config_rows(dev):
    """Device-resident timings of the other BASELINE configs' hot-path rows (each is parity-tested at this size in
    tests/test_gpu_full_size_parity.py): C1 standard image + psf gridding (fp64, 1024^2, cube and continuum), C3 aperture
    image and weight gridding (7-pointing mosaic, 2048^2), C4 degridding predict (fp64, 4096^2).  CUDA events around 5
    back-to-back calls, after 2 warm-up calls; grids are accumulated into (not re-zeroed), which does not change the work."""
    import torch
    from cngi_prototype_b200 import synth, _standard_grid as sg, _aperture_grid as apg, _standard_degrid as sdg
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
    cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(OVERSAMPLING, SUPPORT)).to(dev)

    def ms_of(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def row(ms, n_samples):
        return {"ms": round(ms, 4), "vis_per_s": n_samples / (ms * 1e-3)}

    out = {}
    d = synth.config_c1()
    T = {k: torch.as_tensor(d[k]).to(dev) for k in ("vis", "uvw", "weight", "freq_chan")}
    n = d["weight"].size
    c1 = {"workload": "C1 VLA-like 351 bl x 1000 t x 64 ch x 2 pol = %.1f M samples, 1024^2, S=7, fp64" % (n / 1e6)}
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(1024, d["cell"], chan_mode=mode)
        n_ic = 64 if mode == "cube" else 1
        g = torch.zeros((n_ic, 2, 1024, 1024), dtype=torch.complex128, device=dev)
        pg = torch.zeros((n_ic, 2, 1024, 1024), dtype=torch.float64, device=dev)
        sw, psw = torch.zeros((n_ic, 2), dtype=torch.float64, device=dev), torch.zeros((n_ic, 2), dtype=torch.float64, device=dev)
        c1[mode] = {"image": row(ms_of(lambda: sg.standard_grid(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, False, True,
                                                                grid=g, sum_weight=sw)), n),
                    "psf": row(ms_of(lambda: sg.standard_grid(None, T["uvw"], T["weight"], T["freq_chan"], cgk, gp, True, False,
                                                              grid=pg, sum_weight=psw)), n),
                    "image_and_psf_one_pass": row(ms_of(lambda: sg.standard_grid_image_psf(
                        T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, grid=g, sum_weight=sw, psf_grid=pg,
                        psf_sum_weight=psw, force_fused=True)), n)}
        del g, pg
    out["C1"] = c1
    del T
    d = synth.config_c2(n_time=200, n_chan=64, dtype="f32")
    gcf = synth.make_mosaic_gcf(d["n_baseline"], 64, 2, n_field=7)
    field = synth.mosaic_field_column(200, d["n_baseline"], gcf["field_id"])
    gp = synth.grid_parms_for(2048, d["cell"] * 1.1, chan_mode="continuum")
    gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
    T = {k: torch.as_tensor(d[k]).to(dev) for k in ("vis", "uvw", "weight", "freq_chan")}
    G = {k: torch.as_tensor(v).to(dev) for k, v in gcf.items()}
    fld = torch.as_tensor(field).to(dev)
    common = (T["uvw"], T["weight"], fld, G["cf_baseline_map"], G["cf_chan_map"], G["cf_pol_map"])
    g = torch.zeros((1, 2, 2048, 2048), dtype=torch.complex64, device=dev)
    sw = torch.zeros((1, 2), dtype=torch.float64, device=dev)
    n = d["weight"].size
    out["C3"] = {"workload": "C3 mosaic: 7 pointings, 903 bl x 200 t x 64 ch x 2 pol = %.1f M samples, CF 160^2 (oversampling 10, "
                             "supports 9-15), 2048^2, fp32, continuum" % (n / 1e6),
                 "aperture_image": row(ms_of(lambda: apg._aperture_grid_numpy_wrap(
                     T["vis"], *common, G["conv_kernel"], gcf["weight_support"], G["phase_gradient"], T["freq_chan"], gp, grid=g,
                     sum_weight=sw)), n),
                 "aperture_weight": row(ms_of(lambda: apg._aperture_weight_grid_numpy_wrap(
                     *common, G["weight_conv_kernel"], gcf["weight_support"], G["phase_gradient"], T["freq_chan"], gp, grid=g,
                     sum_weight=sw)), n)}
    del T, G, g
    d = synth.config_c4()
    uvw, freq = torch.as_tensor(d["uvw"]).to(dev), torch.as_tensor(d["freq_chan"]).to(dev)
    gp = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum")
    model = torch.view_as_complex(torch.randn((1, 2, 4096, 4096, 2), dtype=torch.float64, device=dev))
    n = d["weight"].size
    out["C4"] = {"workload": "C4 predict: 351 bl x 1000 t x 64 ch x 2 pol = %.1f M samples from a 4096^2 model grid, S=7, fp64" % (n / 1e6),
                 "degrid": row(ms_of(lambda: sdg._standard_degrid_numpy_wrap(model, uvw, freq, cgk, gp, normalize=True)), n)}
    del model
    torch.cuda.empty_cache()
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    from cngi_prototype_b200 import synth, _lib, imaging
    from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    if world > 1:
        from cngi_prototype_b200 import distributed as _D
        _D.init_nccl(dev)   # NCCL collectives on a high-priority stream (they overlap the gridder instead of queueing behind it)

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
    grid_evs = []

    from types import SimpleNamespace
    from cngi_prototype_b200 import distributed as D
    ops = D.cuda_ops()

    def make_bufs_for(cdt):
        def make_bufs():
            return SimpleNamespace(density=torch.empty((n_ic, 2, a.n_uv, a.n_uv), dtype=torch.float64, device=dev),
                                   dsw=torch.empty((n_ic, 2), dtype=torch.float64, device=dev),
                                   grid=torch.empty((n_ic, 2, a.n_uv, a.n_uv), dtype=cdt, device=dev),
                                   gsw=torch.empty((n_ic, 2), dtype=torch.float64, device=dev))
        return make_bufs

    # the sharding / collective control flow is the one tests/test_distributed_gloo.py exercises on CPU with gloo
    side = torch.cuda.Stream(device=dev, priority=-1) if a.side_stream else None
    pipe = D.ContinuumPipeline(ops, gp, gp_iw, IW_PARMS, cgk_t, make_bufs_for(torch.complex64), side_stream=side,
                               fuse_weights=a.fuse_weights)

    def grid_hook(what):   # CUDA events around the dominant kernel, on the stream it is launched on
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if what == "begin":
            grid_evs.append([ev, None])
        else:
            grid_evs[-1][1] = ev

    def timed(fn, steps, warmup, after=None):
        for _ in range(warmup):
            fn()
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
            fn()
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

    # ---- headline: device-resident steps ------------------------------------------------------------------------
    ms_step, t_begin, t_end = timed(lambda: pipe.step(T, grid_hook=grid_hook), a.steps, max(a.warmup, 3), after=pipe.flush)
    clocks = sampler.window(t_begin, t_end) if rank == 0 else None
    kern_ms = float(np.mean([e0.elapsed_time(e1) for (e0, e1) in grid_evs[-a.steps:] if e1 is not None]))

    # ---- sustained: the same step for >= 3 s (clocks under a long run, not burst) ----------------------------------
    sustained = None
    if not a.no_extras:
        n_sus = max(a.steps, int(3000.0 / max(ms_step, 1e-3)) + 1)
        ms_sus, tb, te = timed(lambda: pipe.step(T), n_sus, 0, after=pipe.flush)
        sustained = {"value": world * n_samples / (ms_sus * 1e-3), "unit": "vis/s", "ms_per_step": ms_sus, "steps": n_sus,
                     "seconds": ms_sus * n_sus * 1e-3, "clocks": sampler.window(tb, te) if rank == 0 else None}

    # ---- parity inputs: one plain (non-pipelined) step, kept for the comparison with the CPU oracle ---------------
    par_gpu = None
    if not a.no_parity and world == 1:
        pb = pipe.bufs[0]
        D.continuum_imaging_step(ops, T, gp, gp_iw, IW_PARMS, cgk_t, pb)
        par_gpu = (pb.grid.cpu().numpy(), pb.gsw.cpu().numpy())
    del pipe
    torch.cuda.empty_cache()

    # ---- the same step at the reference's precision (complex128 / float64) ----------------------------------------
    value_f64 = None
    if not a.no_extras:
        T64 = {"vis": T["vis"].to(torch.complex128), "weight": T["weight"].double(), "uvw": T["uvw"], "freq_chan": T["freq_chan"]}
        pipe64 = D.ContinuumPipeline(ops, gp, gp_iw, IW_PARMS, cgk_t, make_bufs_for(torch.complex128))
        n64 = max(5, min(a.steps, 20))
        ms64, _, _ = timed(lambda: pipe64.step(T64), n64, 3, after=pipe64.flush)
        value_f64 = {"value": world * n_samples / (ms64 * 1e-3), "unit": "vis/s", "ms_per_step": ms64, "steps": n64,
                     "dtype": "f64", "note": "same step and inputs, vis complex128 / weights float64 / grid complex128"}
        del pipe64, T64
        torch.cuda.empty_cache()

    # ---- e2e: the public API on host arrays, from host threads --------------------------------------------------
    e2e = None
    probe = None
    if not a.no_extras:
        probe = pcie_probe(dev, world, dist)
    if not a.no_e2e:
        cell_arcsec = d["cell"] / imaging.ARCSEC_TO_RAD
        api_gp = {"image_size": [a.n_uv, a.n_uv], "cell_size": [cell_arcsec, cell_arcsec], "fft_padding": 1.0,
                  "chan_mode": a.chan_mode}
        ds = {"DATA": H["vis"].numpy(), "UVW": H["uvw"].numpy(), "WEIGHT": H["weight"].numpy(), "chan": H["freq_chan"].numpy()}
        h2d = sum(H[k].numel() * H[k].element_size() for k in H)
        d2h_bytes = [0]
        pend = [None]

        def read_back(g):
            """rank 0 reads the result (the grid of the whole observation lives there); the others only wait for it"""
            if rank == 0:
                grid_h, sw_h = np.asarray(g["GRID"]), np.asarray(g["SUM_WEIGHT"])
                d2h_bytes[0] = grid_h.nbytes + sw_h.nbytes

        def e2e_step():
            """One dataset through the public API.  Results are lazy (like the reference's dask results): the host reads the
            result of dataset k after queueing dataset k+1, so that its D2H overlaps the next H2D (PCIe is full duplex)."""
            if world > 1:
                w = D.make_imaging_weight(ds, IW_PARMS, api_gp)
                g = D.make_grid(w, api_gp, lazy=True)
            else:
                w = imaging.make_imaging_weight(ds, IW_PARMS, api_gp)
                g = imaging.make_grid(w, api_gp, lazy=True)
            if pend[0] is not None:
                read_back(pend[0])
            pend[0] = g

        def e2e_drain():
            if pend[0] is not None:
                read_back(pend[0])
                pend[0] = None

        for _ in range(3):
            e2e_step()
        e2e_drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        e2e_drain()                           # the last result has landed in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ms_e2e = torch.tensor([dt * 1e3 / a.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        ms_e2e = float(ms_e2e.item())
        e2e = {"value": world * n_samples / (ms_e2e * 1e-3), "unit": "vis/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_bytes[0]),
               "h2d_gbs_per_gpu": h2d / (ms_e2e * 1e-3) / 1e9,
               "pcie_frac": (h2d / (ms_e2e * 1e-3) / 1e9 / probe["h2d_gbs_per_gpu"]) if probe else None,
               "pcie_probe": probe,
               "note": "the public API on HOST arrays: %s.make_imaging_weight + make_grid(lazy=True) on numpy views of "
                       "page-locked buffers; the library feeds time chunks H2D on a copy stream under the kernels%s; "
                       "GRID / SUM_WEIGHT of dataset k are read into host memory (rank 0) after dataset k+1 has been queued, "
                       "so the D2H overlaps the next H2D; host wall clock around all calls incl. the last read-back, "
                       "max over ranks; pcie_frac = H2D rate achieved / rate of a plain concurrent H2D probe at this N" % (
                           "distributed" if world > 1 else "imaging",
                           ", NCCL all-reduce(density) + reduce(grid) inside the calls" if world > 1 else "")}

    # ---- config 5 share (cube) ------------------------------------------------------------------------------
    del T
    torch.cuda.empty_cache()
    cube = None
    if not a.no_cube:
        cube = {"config": "BASELINE configs[4] per-GPU share: 8192^2 image, 9830^2 padded grid, %d channels and %.3g samples per GPU, "
                          "fp32, S=7, channel-sharded" % (a.cube_chan_per_gpu, a.cube_samples_per_gpu)}
        cube["time_split_1"] = cube_record(a, dev, rank, world, dist, 1, None)
        if world >= 2 and world % 2 == 0:
            groups = D.make_time_groups(world, 2)
            cube["time_split_2"] = cube_record(a, dev, rank, world, dist, 2, groups)

    if rank != 0:
        sampler.stop()
        if world > 1:
            dist.destroy_process_group()
        return
    sampler.stop()
    configs = config_rows(dev) if (world == 1 and not a.no_configs) else None

    # roofline of the dominant kernel (std_grid_window), algorithmic bytes per SURVEY.md section 8d:
    # n_samples * (8 B vis + 4 B weight) + uvw + grid written once
    peak, peak_src = peaks()
    alg_bytes = n_samples * 12 + a.n_time * d["n_baseline"] * 24 + n_ic * 2 * a.n_uv * a.n_uv * 8
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    tj, tname = traffic_profile()
    traffic, red_sectors = tj.get("std_grid_dram_bytes_per_launch"), tj.get("std_grid_red_sectors_per_launch")
    # the second ceiling BASELINE.json's metric names: reductions into the grid (sectors per launch from the ncu
    # capture of this kernel -- a property of the input and the window, not of timing -- over the live kernel time)
    atomic = None
    std_shape = a.n_time == 500 and a.n_chan == 128 and a.chan_mode == "continuum"
    if red_sectors and std_shape:
        rp = red_peak(n_ic * 2 * a.n_uv * a.n_uv * 8)
        atomic = {"achieved": red_sectors / (kern_ms * 1e-3) / 1e9, "peak": rp, "unit": "Gsector/s",
                  "frac": red_sectors / (kern_ms * 1e-3) / 1e9 / rp, "sectors_per_launch": int(red_sectors),
                  "note": "REDG.E.ADD.F32x2 sectors per launch (ncu l1tex__t_sectors_pipe_lsu_mem_global_op_red, "
                          "profiles/%s) over the live kernel time; peak = cngi_b200_microbench_red, 8 lanes x "
                          "8 B contiguous, footprint = the uv-grid, measured in this run" % tname}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic if std_shape else None, "kernel": tj.get("std_grid_kernel", "std_grid_window_kernel<float,complex,S=7,PP=2>"),
                "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": int(alg_bytes), "peak_source": peak_src,
                "traffic_source": ("ncu --set full, profiles/%s" % tname) if tname else None,
                "note": tj.get("std_grid_note", "not HBM bound: issue slots and the shared-memory pipe limit it; see DESIGN.md section 4.1")}

    # ---- CPU legs (rank 0, N = 1 only): parity against the oracle, then the timed baselines -----------------------
    parity, cb, cb_port = None, None, None
    if world == 1 and (par_gpu is not None or not a.no_cpu_baseline):
        d64 = {"vis": d["vis"].astype(np.complex128), "uvw": d["uvw"], "weight": d["weight"].astype(np.float64),
               "freq_chan": d["freq_chan"], "cell": d["cell"], "n_baseline": d["n_baseline"]}   # exact upcast of the SAME samples
        if not a.no_cpu_baseline or par_gpu is not None:
            cb_port, _, res = cpu_arm(a, 1, 1, a.cpu_sample_times, budget_s=12.0, engine="port", data=d64,
                                      keep_result=par_gpu is not None)
            if par_gpu is not None and res is not None:
                g_ref, s_ref, n_int = res
                if n_int < a.n_time:   # slow host: the oracle took a prefix; grid the same prefix on the GPU
                    Tp = {k: (torch.as_tensor(d[k][:n_int]).to(dev) if k != "freq_chan" else torch.as_tensor(d[k]).to(dev))
                          for k in ("vis", "uvw", "weight", "freq_chan")}
                    pbuf = make_bufs_for(torch.complex64)()
                    D.continuum_imaging_step(ops, Tp, gp, gp_iw, IW_PARMS, cgk_t, pbuf)
                    par_gpu = (pbuf.grid.cpu().numpy(), pbuf.gsw.cpu().numpy())
                    del Tp, pbuf
                g_gpu, s_gpu = par_gpu
                scale = float(np.abs(g_ref).max())
                rel = float(np.abs(g_gpu - g_ref).max() / scale)
                mask_equal = bool(np.array_equal(g_gpu != 0, g_ref != 0))
                sw_rel = float(np.abs(s_gpu - s_ref).max() / np.abs(s_ref).max())
                parity = {"rel_err": rel, "mask_equal": mask_equal, "sum_weight_rel_err": sw_rel,
                          "integrations": int(n_int), "samples": int(n_int * n_samples // a.n_time),
                          "full_size": bool(n_int == a.n_time), "bars": PARITY_BARS,
                          "pass": bool(rel <= PARITY_BARS["rel_err"] and mask_equal and sw_rel <= PARITY_BARS["sum_weight_rel_err"]),
                          "oracle": "oracle/cngi_oracle.c multi-threaded (fp64) on the exact upcast of the GPU arm's fp32 samples: "
                                    "density grid -> Briggs -> weight degrid -> gridding; rel_err = max|gpu - cpu| / max|cpu| over "
                                    "the uv-grid, mask = which cells are non-zero"}
                del g_ref, g_gpu
        if not a.no_cpu_baseline:
            engine = pick_engine(a.cpu_engine)
            if engine == "reference":
                cb, _, _ = cpu_arm(a, 1, 1, a.cpu_sample_times, budget_s=24.0, engine="reference", data=d64)
            else:
                cb, cb_port = cb_port, None

    line = {"metric": METRIC, "value": world * n_samples / (ms_step * 1e-3), "unit": "vis/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "samples_per_gpu": int(n_samples), "valid_samples_per_gpu": n_valid,
                       "l2": "inputs (1.4 GB/step) are larger than L2 (126 MB), no explicit flush",
                       "parallelism": ("time-sharded x%d, NCCL all-reduce(density plane 0) + reduce(grid), overlapped across steps" % world) if world > 1 else "single GPU"},
            "vis_tap_per_s": world * n_samples * SUPPORT * SUPPORT / (ms_step * 1e-3),
            "gridding_kernel_vis_per_s": n_samples / (kern_ms * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": 5 * a.steps,
            "gpu_launches_note": "per step of the timed (device-resident) region: iw_grid, iw_sumsq, iw_briggs_finalize, iw_degrid, "
                                 "std_grid_window; the 128-thread uv_scale table kernel in front of A2 / A4 / A1 and the memsets are not counted",
            "roofline": roofline, "atomic_roofline": atomic, "parity": parity,
            "sustained": sustained, "value_f64": value_f64, "cube": cube, "configs": configs,
            "cpu_baseline": cb, "cpu_baseline_port": cb_port}
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
