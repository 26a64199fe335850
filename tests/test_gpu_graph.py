"""Graph-level boundary (cngi_prototype_b200/_graph.py: _graph_standard_grid / _graph_standard_degrid /
_graph_aperture_grid with the reference's names, _standard_grid.py:23,381 and _aperture_grid.py:25) against the oracle's
per-chunk operators applied to the whole arrays: chunking along time / baseline / chan must not change the result
beyond summation order."""
import numpy as np
import pytest

from _util import rel_err, same_support

pytestmark = pytest.mark.gpu
SEL = {"data_group_in": {"data": "DATA", "uvw": "UVW", "imaging_weight": "IMAGING_WEIGHT"}}


def _dataset(seed=33):
    from cngi_prototype_b200 import synth
    d = synth.make_vis_set(7, 17, 5, 2, 1.0e9, 1.1e9, 300.0, 90.0, seed=seed)
    ds = {"DATA": d["vis"], "UVW": d["uvw"], "IMAGING_WEIGHT": d["weight"], "chan": d["freq_chan"],
          "chunks": {"time": 5, "baseline": 8, "chan": 2}}
    return d, ds


@pytest.mark.parametrize("mode", ["cube", "continuum"])
def test_graph_standard_grid_and_weights(oracle, mode):
    from cngi_prototype_b200 import synth, _graph
    d, ds = _dataset()
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    gp = synth.grid_parms_for(128, d["cell"], chan_mode=mode)
    g, s = _graph._graph_standard_grid(ds, cgk, gp, SEL)
    gr, sr = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gp)
    assert g.shape == np.moveaxis(gr, (0, 1), (2, 3)).shape
    assert same_support(g, np.moveaxis(gr, (0, 1), (2, 3))) and rel_err(g, np.moveaxis(gr, (0, 1), (2, 3))) < 1e-12
    assert rel_err(s, sr) < 1e-12
    gpp = dict(gp, do_psf=True, complex_grid=False)
    g, s = _graph._graph_standard_grid(ds, cgk, gpp, SEL)
    gr, sr = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk, gpp)
    assert not np.iscomplexobj(g) and rel_err(g, np.moveaxis(gr, (0, 1), (2, 3))) < 1e-12 and rel_err(s, sr) < 1e-12
    # imaging weights: density graph -> briggs factors -> degrid graph (make_imaging_weight.py:153-161,244)
    gpw = synth.grid_parms_for(128, d["cell"], chan_mode=mode, support=1, oversampling=0, do_psf=True, complex_grid=False,
                               do_imaging_weight=True)
    rho, sw = _graph._graph_standard_grid(ds, np.ones(1), gpw, SEL)
    rr, swr = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gpw)
    assert rel_err(rho, np.moveaxis(rr, (0, 1), (2, 3))) < 1e-12 and rel_err(sw, swr) < 1e-12
    bf = oracle._calculate_briggs_parms(rr, swr, dict(weighting="briggs", robust=0.5))
    iw = _graph._graph_standard_degrid(ds, np.moveaxis(rr, (0, 1), (2, 3)), bf, None, gpw, SEL)
    iwr = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rr, (0, 1), (2, 3)), d["uvw"], d["weight"], bf,
                                                            d["freq_chan"], gpw)
    assert np.array_equal(np.isnan(iw), np.isnan(iwr))
    m = np.isfinite(iwr)
    assert np.max(np.abs(iw[m] - iwr[m])) <= 1e-12 * np.max(np.abs(iwr[m]))


@pytest.mark.parametrize("mode", ["cube", "continuum"])
def test_graph_aperture_grid(oracle, mode):
    from cngi_prototype_b200 import synth, _graph
    d, ds = _dataset(seed=34)
    n_b = d["n_baseline"]
    gcf = synth.make_mosaic_gcf(n_b, 5, 2, n_field=3, n_cf_baseline=2, n_cf_chan=2, oversampling=(4, 5), max_support=(9, 9), seed=4)
    fld = synth.mosaic_field_column(17, n_b, gcf["field_id"], frac_unset=0.05)
    ds["FIELD_ID"] = fld
    gds = dict(CONV_KERNEL=gcf["conv_kernel"], WEIGHT_CONV_KERNEL=gcf["weight_conv_kernel"], SUPPORT=gcf["weight_support"],
               PHASE_GRADIENT=gcf["phase_gradient"], CF_BASELINE_MAP=gcf["cf_baseline_map"], CF_CHAN_MAP=gcf["cf_chan_map"],
               CF_POL_MAP=gcf["cf_pol_map"], field_id=gcf["field_id"], oversampling=np.asarray(gcf["oversampling"]))
    gp = synth.grid_parms_for(112, d["cell"] * 1.2, chan_mode=mode)
    gp["oversampling"] = np.asarray(gcf["oversampling"])
    ogp = dict(gp, field_id=gcf["field_id"])
    common = (d["uvw"], d["weight"], fld, gcf["cf_baseline_map"], gcf["cf_chan_map"], gcf["cf_pol_map"])
    tail = (gcf["weight_support"], gcf["phase_gradient"], d["freq_chan"])
    cases = [(dict(grid_weights=False, do_psf=False), oracle._aperture_grid_numpy_wrap(d["vis"], *common, gcf["conv_kernel"], *tail, ogp)),
             (dict(grid_weights=False, do_psf=True), oracle._aperture_psf_grid_numpy_wrap(*common, gcf["conv_kernel"], *tail, dict(ogp, do_psf=True))),
             (dict(grid_weights=True, do_psf=False), oracle._aperture_weight_grid_numpy_wrap(*common, gcf["weight_conv_kernel"], *tail, ogp))]
    for flags, (gr, sr) in cases:
        g, s = _graph._graph_aperture_grid(ds, gds, dict(gp, **flags), SEL)
        ga = np.moveaxis(gr, (0, 1), (2, 3))
        assert g.shape == ga.shape and same_support(g, ga), flags
        assert rel_err(g, ga) < 1e-12 and rel_err(s, sr) < 1e-12, flags
