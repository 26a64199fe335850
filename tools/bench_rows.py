"""Times every hot-path row of SURVEY.md section 8 at (or near) its BASELINE.json configuration on one B200 and, for a
bounded sample, the CPU oracle port beside it.  Development / documentation tool (the contract benchmark is bench.py).

    python tools/bench_rows.py [--quick] > profiles/r01_rows.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_grid, _imaging_weight, _aperture_grid, _standard_degrid, _fft  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import (_create_prolate_spheroidal_kernel_1D,  # noqa: E402
                                                                   correcting_function_1D)
from oracle import oracle as O  # noqa: E402  (CPU column only)


def gpu_ms(fn, n=10, warm=3, reps=3):
    """ms per call: `n` back-to-back asynchronous calls between two events (so the host-side cost of one call hides behind
    the previous call's kernels, as in a pipeline), best of `reps`.  Rows whose kernels are shorter than the Python
    wrapper (~0.2 ms of argument marshalling) stay wrapper-bound here; profiles/r01_launch_shares.md has their kernel
    times."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    single = []
    for _ in range(5):   # one call at a time (what round 1's first table reported); the lower of the two is kept
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        single.append(e0.elapsed_time(e1))
    return min(best, float(np.median(single)))


def cpu_s(fn):
    fn()   # warm (page faults, thread start)
    t = time.perf_counter()
    fn()
    return time.perf_counter() - t


def dev(x):
    return torch.as_tensor(x).cuda()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    q = a.quick
    rows = []
    threads = min(os.cpu_count() or 1, 32)
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)

    def add(row, config, n_samples, ms, cpu_samples=None, cpu_sec=None, **extra):
        r = dict(row=row, config=config, samples=int(n_samples), gpu_ms=round(ms, 3),
                 gpu_gvis_per_s=round(n_samples / ms / 1e6, 2))
        if cpu_sec:
            r.update(cpu_mvis_per_s=round(cpu_samples / cpu_sec / 1e6, 2), cpu_threads=threads,
                     cpu_sample_samples=int(cpu_samples))
        r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    # ---- C1: VLA-like fp64, 1024^2, cube and continuum: A1 image + psf -----------------------------------------------
    d = synth.config_c1(n_time=200 if q else 1000)
    ds = synth.config_c1(n_time=50)
    T = {k: dev(d[k]) for k in ("vis", "uvw", "weight", "freq_chan")}
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(1024, d["cell"], chan_mode=mode)
        gpp = dict(gp, do_psf=True, complex_grid=False)
        n_ic = 64 if mode == "cube" else 1
        grid = torch.zeros((n_ic, 2, 1024, 1024), dtype=torch.complex128, device="cuda")
        sw = torch.zeros((n_ic, 2), dtype=torch.float64, device="cuda")
        ms = gpu_ms(lambda: _standard_grid.standard_grid(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, False, True,
                                                         grid=grid, sum_weight=sw))
        c = cpu_s(lambda: O._standard_grid_numpy_wrap(ds["vis"], ds["uvw"], ds["weight"], ds["freq_chan"], cgk, gp,
                                                      n_threads=threads))
        add("A1 image", "C1 VLA-like fp64 1024^2 %s" % mode, d["weight"].size, ms, ds["weight"].size, c)
        pgrid = torch.zeros((n_ic, 2, 1024, 1024), dtype=torch.float64, device="cuda")
        ms = gpu_ms(lambda: _standard_grid.standard_grid(None, T["uvw"], T["weight"], T["freq_chan"], cgk, gpp, True, False,
                                                         grid=pgrid, sum_weight=sw))
        c = cpu_s(lambda: O._standard_grid_psf_numpy_wrap(ds["uvw"], ds["weight"], ds["freq_chan"], cgk, gpp, n_threads=threads))
        add("A1 psf", "C1 VLA-like fp64 1024^2 %s" % mode, d["weight"].size, ms, ds["weight"].size, c)
        if mode == "cube":   # A9/A10 on the same cube: 128 planes of 1228^2 would be the padded case; here the grid as gridded
            cu, cv = correcting_function_1D([1024, 1024], [854, 854])
            ms = gpu_ms(lambda: _fft.grid_to_image(grid, [854, 854], sum_weight=sw, corr_u=cu, corr_v=cv))
            gh = grid[:8].cpu().numpy()
            c = cpu_s(lambda: O.correct_image(O.grid_to_uncorrected_image(gh, np.array([854, 854])), np.ones((8, 2)),
                                              O._remove_padding(O._create_prolate_spheroidal_image_2D([1024, 1024]),
                                                                np.array([854, 854]))))
            add("A9+A10 grid->image", "128 planes 1024^2 -> 854^2 fp64 (cuFFT Z2Z + post kernel)", 128, ms,
                extra_note="samples column = planes; CPU: numpy.fft on 16 planes, %.2f s" % c,
                gpu_planes_per_s=round(128 / ms * 1e3, 1), cpu_planes_per_s=round(16 / c, 2))
        del grid, pgrid
    del T
    torch.cuda.empty_cache()

    # ---- C2: ALMA-like fp32, 4096^2 continuum: A2, A3, A4, A1 ----------------------------------------------------------
    d = synth.config_c2(n_time=100 if q else 500, dtype="f32")
    ds = synth.config_c2(n_time=25, dtype="f64")
    T = {k: dev(d[k]) for k in ("vis", "uvw", "weight", "freq_chan")}
    gp = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum")
    gpw = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                               complex_grid=False, do_imaging_weight=True)
    rho = torch.zeros((1, 2, 4096, 4096), dtype=torch.float64, device="cuda")
    rsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
    ms = gpu_ms(lambda: _imaging_weight.imaging_weight_grid(T["uvw"], T["weight"], T["freq_chan"], gpw, grid=rho, sum_weight=rsw))
    c = cpu_s(lambda: O._standard_grid_psf_numpy_wrap(ds["uvw"], ds["weight"], ds["freq_chan"], np.ones(1), gpw, n_threads=threads))
    add("A2 density grid", "C2 ALMA-like fp32 4096^2 continuum", d["weight"].size, ms, ds["weight"].size, c)
    rho.zero_(), rsw.zero_()
    _imaging_weight.imaging_weight_grid(T["uvw"], T["weight"], T["freq_chan"], gpw, grid=rho, sum_weight=rsw)
    ms = gpu_ms(lambda: _imaging_weight.calculate_briggs_parms(rho, rsw, dict(weighting="briggs", robust=0.5)))
    add("A3 briggs factors", "2 planes of 4096^2 fp64", 2 * 4096 * 4096, ms, extra_note="samples column = grid cells")
    bf = _imaging_weight.calculate_briggs_parms(rho, rsw, dict(weighting="briggs", robust=0.5))
    ms = gpu_ms(lambda: _imaging_weight._standard_imaging_weight_degrid_numpy_wrap(rho, T["uvw"], T["weight"], bf, T["freq_chan"],
                                                                                  gpw, kernel_side_layout=True))
    rho_h, bf_h = np.moveaxis(rho.cpu().numpy(), (0, 1), (2, 3)), bf.cpu().numpy()
    c = cpu_s(lambda: O._standard_imaging_weight_degrid_numpy_wrap(rho_h, ds["uvw"], ds["weight"], bf_h, ds["freq_chan"], gpw))
    add("A4 weight degrid", "C2 ALMA-like fp32 4096^2 continuum", d["weight"].size, ms, ds["weight"].size, c,
        cpu_note="single thread (the oracle degrid is not threaded)")
    grid = torch.zeros((1, 2, 4096, 4096), dtype=torch.complex64, device="cuda")
    gsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
    for algo, name in ((4, "window"), (2, "track"), (1, "naive")):
        ms = gpu_ms(lambda: _standard_grid.standard_grid(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, False, True,
                                                         grid=grid, sum_weight=gsw, algorithm=algo))
        add("A1 image (%s kernel)" % name, "C2 ALMA-like fp32 4096^2 continuum", d["weight"].size, ms)
    c = cpu_s(lambda: O._standard_grid_numpy_wrap(ds["vis"], ds["uvw"], ds["weight"], ds["freq_chan"], cgk, gp, n_threads=threads))
    rows[-3].update(cpu_mvis_per_s=round(ds["weight"].size / c / 1e6, 2), cpu_threads=threads)
    # A9 on the 4096^2 continuum grid, default 1.2x padding would be 4915 (5 * 983): time both
    cu, cv = correcting_function_1D([4096, 4096], [3412, 3412])
    ms = gpu_ms(lambda: _fft.grid_to_image(grid, [3412, 3412], sum_weight=gsw, corr_u=cu, corr_v=cv))
    add("A9+A10 grid->image", "2 planes 4096^2 -> 3412^2 fp32 (cuFFT C2C + post kernel)", 2, ms,
        gpu_planes_per_s=round(2 / ms * 1e3, 1), extra_note="samples column = planes")
    g4915 = torch.zeros((1, 2, 4915, 4915), dtype=torch.complex64, device="cuda")
    cu, cv = correcting_function_1D([4915, 4915], [4096, 4096])
    ms = gpu_ms(lambda: _fft.grid_to_image(g4915, [4096, 4096], sum_weight=gsw, corr_u=cu, corr_v=cv))
    add("A9+A10 grid->image", "2 planes 4915^2 (=5*983, odd) -> 4096^2 fp32", 2, ms, gpu_planes_per_s=round(2 / ms * 1e3, 1),
        extra_note="samples column = planes; default fft_padding 1.2 gives a 983-prime size")
    del g4915, rho
    # ---- C4: degrid predict on the C2 set (4096^2, S=7) -------------------------------------------------------------------
    model = torch.randn((1, 2, 4096, 4096), dtype=torch.complex64, device="cuda")
    ms = gpu_ms(lambda: _standard_degrid._standard_degrid_numpy_wrap(model, T["uvw"], T["freq_chan"], cgk, gp, normalize=True))
    mh = model.cpu().numpy()
    c = cpu_s(lambda: O._standard_degrid_numpy_wrap(mh, ds["uvw"], ds["freq_chan"], cgk, gp, normalize=True))
    add("A7 degrid predict", "C2 geometry fp32 4096^2 S=7 (C4-like)", d["weight"].size, ms, ds["weight"].size, c,
        cpu_note="single thread")
    del model, grid, T
    torch.cuda.empty_cache()

    # ---- C3: mosaic aperture gridding, 7 pointings, CF 160x160, 2048^2 ----------------------------------------------------
    d = synth.config_c2(n_time=50 if q else 200, n_chan=64, dtype="f32")
    ds = synth.config_c2(n_time=4, n_chan=64, dtype="f64")
    gcf = synth.make_mosaic_gcf(d["n_baseline"], 64, 2, n_field=7)
    for dd in (d, ds):
        dd["field"] = synth.mosaic_field_column(dd["uvw"].shape[0], dd["n_baseline"], gcf["field_id"])
    gp = synth.grid_parms_for(2048, d["cell"] * 1.1, chan_mode="continuum")
    gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
    T = {k: dev(d[k]) for k in ("vis", "uvw", "weight", "freq_chan", "field")}
    G = {k: dev(v) for k, v in gcf.items()}
    grid = torch.zeros((1, 2, 2048, 2048), dtype=torch.complex64, device="cuda")
    gsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
    common = lambda X, GG: (X["uvw"], X["weight"], X["field"], GG["cf_baseline_map"], GG["cf_chan_map"], GG["cf_pol_map"])
    ms = gpu_ms(lambda: _aperture_grid._aperture_grid_numpy_wrap(T["vis"], *common(T, G), G["conv_kernel"], gcf["weight_support"],
                                                                 G["phase_gradient"], T["freq_chan"], gp, grid=grid, sum_weight=gsw))
    c = cpu_s(lambda: O._aperture_grid_numpy_wrap(ds["vis"], *common(ds, gcf), gcf["conv_kernel"], gcf["weight_support"],
                                                  gcf["phase_gradient"], ds["freq_chan"], gp))
    sup = gcf["weight_support"][..., 0]
    add("A5 aperture image", "C3 mosaic 7 fields, CF 160^2 os 10, supports %s, 2048^2 fp32 continuum" % sorted(set(sup.ravel().tolist())),
        d["weight"].size, ms, ds["weight"].size, c, cpu_note="single thread")
    ms = gpu_ms(lambda: _aperture_grid._aperture_weight_grid_numpy_wrap(*common(T, G), G["weight_conv_kernel"], gcf["weight_support"],
                                                                        G["phase_gradient"], T["freq_chan"], gp, grid=grid, sum_weight=gsw))
    c = cpu_s(lambda: O._aperture_weight_grid_numpy_wrap(*common(ds, gcf), gcf["weight_conv_kernel"], gcf["weight_support"],
                                                         gcf["phase_gradient"], ds["freq_chan"], gp))
    add("A6 aperture weight grid", "C3 mosaic, 2048^2 fp32 continuum", d["weight"].size, ms, ds["weight"].size, c,
        cpu_note="single thread")
    del T, G, grid
    torch.cuda.empty_cache()

    # ---- section 8(f) rows: N1 fused image+psf, N2 GCF on the device, N3 direction_rotate ------------------------------
    from cngi_prototype_b200 import direction_rotate as dr, make_gridding_convolution_function as mg
    d = synth.config_c2(n_time=100 if q else 500, dtype="f32")
    ds = synth.config_c2(n_time=25, dtype="f64")
    T = {k: dev(d[k]) for k in ("vis", "uvw", "weight", "freq_chan")}
    gp = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum")
    g = torch.zeros((1, 2, 4096, 4096), dtype=torch.complex64, device="cuda")
    pg = torch.zeros((1, 2, 4096, 4096), dtype=torch.float32, device="cuda")
    sw, psw = torch.zeros((1, 2), dtype=torch.float64, device="cuda"), torch.zeros((1, 2), dtype=torch.float64, device="cuda")
    ms_i = gpu_ms(lambda: _standard_grid.standard_grid(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, False, True, grid=g, sum_weight=sw))
    ms_p = gpu_ms(lambda: _standard_grid.standard_grid(None, T["uvw"], T["weight"], T["freq_chan"], cgk, gp, True, False, grid=pg, sum_weight=psw))
    ms = gpu_ms(lambda: _standard_grid.standard_grid_image_psf(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, grid=g,
                                                               sum_weight=sw, psf_grid=pg, psf_sum_weight=psw))
    add("N1 fused image+psf pass", "C2 ALMA-like fp32 4096^2 continuum", d["weight"].size, ms,
        two_passes_ms=round(ms_i + ms_p, 3), image_ms=round(ms_i, 3), psf_ms=round(ms_p, 3))
    del g, pg
    n_t, n_b = d["uvw"].shape[0], d["uvw"].shape[1]
    ids = np.arange(7)
    dirs = np.stack([1.0 + 4e-4 * np.cos(ids), 0.5 + 4e-4 * np.sin(ids)], 1)
    field = np.repeat((np.arange(n_t) % 7)[:, None], n_b, 1).astype(np.int64)
    R, P, rid = dr.calc_rotation_mats(field, ids, dirs, dict(new_phase_center=[1.0, 0.5]))
    Fd, Rd, Pd, rd = dev(field), dev(R), dev(P), dev(rid)
    ms = gpu_ms(lambda: dr.rotate_chunk(T["vis"], T["uvw"], Fd, T["freq_chan"], Rd, Pd, rd, True, False))
    fs = field[:25]
    c = cpu_s(lambda: O.apply_phasor(ds["vis"], O.apply_rotation_matrix(ds["uvw"], fs, R, rid), fs, ds["freq_chan"], P, rid, True, False))
    add("N3 direction_rotate", "C2/C3 sample shape complex64, 7 fields (through the Python mirror; kernels 0.32 ms, "
        "profiles/r01_direction_rotate_phasor_f32.txt)", d["weight"].size, ms, ds["weight"].size, c, cpu_note="numpy, single thread")
    del T
    torch.cuda.empty_cache()
    n_ant = 43
    a1, a2 = np.triu_indices(n_ant, 1)
    for name, types, dishes, blocks in (("one dish type", np.zeros(n_ant, dtype=int), [10.7], [0.75]),
                                        ("12 m + 7 m array", (np.arange(n_ant) % 4 == 0).astype(int), [10.7, 6.25], [0.75, 0.75])):
        gparms = dict(function="casa_airy", list_dish_diameters=np.array(dishes), list_blockage_diameters=np.array(blocks),
                      unique_ant_indx=types, basline_ant=np.stack([a1, a2], 1), freq_chan=np.linspace(345e9, 347e9, 128),
                      pol=np.array([0, 1]), field_phase_dir=dirs, phase_center=np.array([1.0, 0.5]), oversampling=[10, 10],
                      max_support=[15, 15])
        grid_parms = dict(image_size=np.array([2048, 2048]), image_size_padded=np.array([2048, 2048]),
                          cell_size=np.array([-0.02, 0.02]) * np.pi / (180 * 3600))
        out = mg.make_gridding_convolution_function(gparms, grid_parms)
        ms = gpu_ms(lambda: mg.make_gridding_convolution_function(gparms, grid_parms), n=3, warm=1, reps=2)
        c = cpu_s(lambda: O.make_gridding_convolution_function(gparms, grid_parms)) if not q else None
        items = int(out["CONV_KERNEL"].shape[0] * out["CONV_KERNEL"].shape[1])
        add("N2 make_gridding_convolution_function", "C3: n_pad 2048^2, os 10, max_support 15, 7 fields, %s" % name, items, ms,
            extra_note="samples column = (antenna-type pair, PB frequency) items; each is two 2048^2 Z2Z FFTs",
            cpu_oracle_s=None if c is None else round(c, 2), supports=sorted(set(out["SUPPORT"].cpu().numpy().ravel().tolist())))
    print(json.dumps({"rows": rows}))


if __name__ == "__main__":
    main()
