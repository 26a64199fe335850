"""Generates tests/golden/*.npz by running the UNMODIFIED reference (this container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden.py

Every fixture stores the inputs and the outputs of the reference's own numba wrappers
(/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py:123,180,443 and
_aperture_grid.py:146,294,333, _gridding_convolutional_kernels.py:35,151), loaded by
oracle/ref_loader.py.  The fixtures pin the oracle (tests/test_oracle_golden.py, bit-exact)
and anchor the CUDA parity tests (tests/test_gpu_*.py).  /root/reference is not needed to
RUN the tests, only to regenerate these files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402
from cngi_prototype_b200 import synth  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-32s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def gp_arrays(gp):
    return dict(gp_chan_mode=np.array(gp["chan_mode"]), gp_image_size_padded=gp["image_size_padded"],
                gp_cell_size=gp["cell_size"], gp_oversampling=np.asarray(gp["oversampling"]),
                gp_support=np.array(gp.get("support", 0)), gp_do_psf=np.array(gp["do_psf"]),
                gp_complex_grid=np.array(gp["complex_grid"]),
                gp_do_imaging_weight=np.array(gp["do_imaging_weight"]))


def main():
    sg, ag, ck = ref_loader.load()

    # ---- A8 tables -----------------------------------------------------------------------
    cgk = ck._create_prolate_spheroidal_kernel_1D(100, 7)
    _, img12 = ck._create_prolate_spheroidal_kernel(100, 7, np.array([12, 12]))
    _, img_odd = ck._create_prolate_spheroidal_kernel(100, 7, np.array([15, 13]))
    cgk_5 = ck._create_prolate_spheroidal_kernel_1D(50, 5)
    save("ps_tables", cgk_1D_os100_s7=cgk, corr_image_12x12=img12, corr_image_15x13=img_odd,
         cgk_1D_os50_s5=cgk_5)

    # ---- single-sample known answer --------------------------------
    gp = dict(chan_mode="cube", image_size_padded=np.array([64, 64]), cell_size=np.array([-1 / 64, 1 / 64]),
              oversampling=100, support=7, complex_grid=True, do_psf=False, do_imaging_weight=False)
    vis = np.full((1, 1, 1, 1), 2 - 1j, dtype=np.complex128)
    uvw = np.array([[[10.3, 4.7, 0.0]]])
    w = np.full((1, 1, 1, 1), 0.5)
    freq = np.array([299792458.0])
    g, s = sg._standard_grid_numpy_wrap(vis, uvw, w, freq, cgk, gp)
    save("std_single_sample", vis=vis, uvw=uvw, weight=w, freq_chan=freq, cgk_1D=cgk, grid=g, sum_weight=s,
         **gp_arrays(gp))

    # ---- half-way / edge cases: uv_scale = (+1,-1) cell per metre ----------------------------
    pts = np.array([[10.5, 4.5], [-10.5, -4.5], [3.005, -7.995], [0.0, 0.0], [28.4, 0.2], [28.6, 0.2],
                    [-28.49, 0.0], [-28.51, 0.0], [5.0, 28.49], [5.0, -28.6], [np.nan, 1.0], [1.0, np.nan],
                    [-31.2, 31.2], [12.25, -12.75], [0.495, 0.505], [-0.505, -0.495]])
    uvw = np.zeros((len(pts), 1, 3))
    uvw[:, 0, :2] = pts
    rng = np.random.default_rng(5)
    vis = (rng.standard_normal((len(pts), 1, 1, 2)) + 1j * rng.standard_normal((len(pts), 1, 1, 2)))
    w = rng.uniform(0.5, 1.5, size=(len(pts), 1, 1, 2))
    g, s = sg._standard_grid_numpy_wrap(vis, uvw, w, freq, cgk, gp)
    save("std_halfway_edges", vis=vis, uvw=uvw, weight=w, freq_chan=freq, cgk_1D=cgk, grid=g, sum_weight=s,
         **gp_arrays(gp))

    # ---- standard gridder, image + psf, cube + continuum, non-square odd grid ------------------
    d = synth.make_vis_set(6, 14, 3, 2, 1.0e9, 1.1e9, 300.0, 120.0, seed=21)
    for mode in ("cube", "continuum"):
        for n_uv, tag, cellfac in (((64, 64), "sq", 1.0), ((61, 75), "odd", 1.25)):
            gp = synth.grid_parms_for(64, d["cell"] * cellfac, chan_mode=mode)
            gp["image_size_padded"] = np.array(n_uv, dtype=np.int64)
            g, s = sg._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
            gpp = dict(gp, do_psf=True, complex_grid=False)
            gpsf, spsf = sg._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk, gpp)
            save("std_%s_%s" % (mode, tag), vis=d["vis"], uvw=d["uvw"], weight=d["weight"],
                 freq_chan=d["freq_chan"], cgk_1D=cgk, grid=g, sum_weight=s, psf_grid=gpsf,
                 psf_sum_weight=spsf, **gp_arrays(gp))

    # ---- support 5 / oversampling 50, single pol ---------------------------------------------
    d1 = synth.make_vis_set(5, 10, 2, 1, 1.0e9, 1.05e9, 300.0, 120.0, seed=22)
    gp = synth.grid_parms_for(48, d1["cell"], chan_mode="cube", support=5, oversampling=50)
    g, s = sg._standard_grid_numpy_wrap(d1["vis"], d1["uvw"], d1["weight"], d1["freq_chan"], cgk_5, gp)
    save("std_cube_s5_1pol", vis=d1["vis"], uvw=d1["uvw"], weight=d1["weight"], freq_chan=d1["freq_chan"],
         cgk_1D=cgk_5, grid=g, sum_weight=s, **gp_arrays(gp))

    # ---- imaging weights: density grid (A2), briggs factors (A3), degrid (A4) -----------------
    for dd, tag in ((d, "2pol"), (d1, "1pol")):
        for mode in ("cube", "continuum"):
            n = 64
            gp = synth.grid_parms_for(n, dd["cell"], chan_mode=mode, support=1, oversampling=0, do_psf=True,
                                      complex_grid=False, do_imaging_weight=True)
            rho, sw = sg._standard_grid_psf_numpy_wrap(dd["uvw"], dd["weight"], dd["freq_chan"], np.ones(1), gp)
            # calculate_briggs_parms (make_imaging_weight.py:198-213), robust 0.5, on the API-side layout
            rho_api = np.moveaxis(rho, (0, 1), (2, 3))
            bf = np.ones((2,) + sw.shape)
            bf[0] = np.square(5.0 * 10.0 ** (-0.5)) / (np.sum(rho_api ** 2, axis=(0, 1)) / sw)
            iw = sg._standard_imaging_weight_degrid_numpy_wrap(rho_api, dd["uvw"], dd["weight"], bf,
                                                              dd["freq_chan"], gp)
            save("iw_%s_%s" % (mode, tag), uvw=dd["uvw"], weight=dd["weight"], freq_chan=dd["freq_chan"],
                 density=rho, sum_weight=sw, briggs_factors=bf, imaging_weight=iw, **gp_arrays(gp))

    # ---- aperture gridders (A5/A6) -------------------------------------------------------------
    da = synth.make_vis_set(6, 12, 4, 2, 1.0e9, 1.1e9, 300.0, 120.0, seed=23)
    gcf = synth.make_mosaic_gcf(da["n_baseline"], 4, 2, n_field=3, n_cf_baseline=2, n_cf_chan=2,
                                oversampling=(4, 5), max_support=(9, 9), seed=3)
    fld = synth.mosaic_field_column(12, da["n_baseline"], gcf["field_id"], frac_unset=0.05)
    for mode in ("cube", "continuum"):
        gp = synth.grid_parms_for(96, da["cell"] * 1.2, chan_mode=mode)
        gp["oversampling"] = gcf["oversampling"]
        gp["field_id"] = gcf["field_id"]
        common = (da["uvw"], da["weight"], fld, gcf["cf_baseline_map"], gcf["cf_chan_map"], gcf["cf_pol_map"])
        g, s = ag._aperture_grid_numpy_wrap(da["vis"], *common, gcf["conv_kernel"], gcf["weight_support"],
                                            gcf["phase_gradient"], da["freq_chan"], gp)
        gp_psf = dict(gp, do_psf=True)
        gpsf, spsf = ag._aperture_psf_grid_numpy_wrap(*common, gcf["conv_kernel"], gcf["weight_support"],
                                                      gcf["phase_gradient"], da["freq_chan"], gp_psf)
        gw, sw = ag._aperture_weight_grid_numpy_wrap(*common, gcf["weight_conv_kernel"], gcf["weight_support"],
                                                     gcf["phase_gradient"], da["freq_chan"], gp)
        arrs = gp_arrays(gp)
        save("aperture_%s" % mode, vis=da["vis"], uvw=da["uvw"], weight=da["weight"], freq_chan=da["freq_chan"],
             field=fld, grid=g, sum_weight=s, psf_grid=gpsf, psf_sum_weight=spsf, weight_grid=gw,
             weight_sum_weight=sw, gp_field_id=gcf["field_id"],
             **{"gcf_" + k: v for k, v in gcf.items()}, **arrs)


if __name__ == "__main__":
    main()
