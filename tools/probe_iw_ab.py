"""A/B of the imaging-weight kernels (development tool): the round-1 kernels (CNGI_IW_GRID_OLD=1 / CNGI_IW_DEGRID_MLP=1)
against the fast ones on config C2 -- density and sum_weight to 1e-12, imaging weights bit for bit -- with device timings."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _imaging_weight  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402


def run(prec, mode, n_pol, n_time, n_chan=128):
    d = synth.config_c2(n_time=n_time, n_chan=n_chan, dtype=prec)
    T = {k: torch.as_tensor(d[k]).cuda() for k in ("uvw", "weight", "freq_chan")}
    if n_pol == 1:
        T["weight"] = T["weight"][..., :1].contiguous()
    w = T["weight"]
    w[3, 5, 2, 0] = float("nan")
    w[4, 7, 1, :] = 0.0
    T["uvw"][6, 11, 0] = float("nan")
    n_ic = n_chan if mode == "cube" else 1
    gpw = synth.grid_parms_for(4096 if mode == "continuum" else 1024, d["cell"], chan_mode=mode, support=1, oversampling=0,
                               do_psf=True, complex_grid=False, do_imaging_weight=True)
    n = int(gpw["image_size_padded"][0])
    out = {}
    res = {}
    for tag, env in (("old", "1"), ("new", "0")):
        os.environ["CNGI_IW_GRID_OLD"] = env
        os.environ["CNGI_IW_DEGRID_MLP"] = env
        for fpo in ((True, False) if n_pol == 2 else (False,)):
            rho = torch.zeros((n_ic, n_pol, n, n), dtype=torch.float64, device="cuda")
            rsw = torch.zeros((n_ic, n_pol), dtype=torch.float64, device="cuda")
            f = lambda: _imaging_weight.imaging_weight_grid(T["uvw"], w, T["freq_chan"], gpw, grid=rho, sum_weight=rsw,  # noqa: E731
                                                            first_pol_only=fpo)
            f()
            torch.cuda.synchronize()
            res[(tag, "rho", fpo)] = rho.clone(), rsw.clone()
            out["%s iw_grid fpo=%d ms" % (tag, fpo)] = round(timeit(f)[0], 4)
        rho, rsw = res[(tag, "rho", False)]
        bf = _imaging_weight.calculate_briggs_parms(rho, rsw, dict(weighting="briggs", robust=0.5))
        for shared in ((False, True) if n_pol == 2 else (False,)):
            r = rho[:, :1].expand(-1, n_pol, -1, -1) if shared else rho
            f = lambda: _imaging_weight._standard_imaging_weight_degrid_numpy_wrap(  # noqa: E731
                r, T["uvw"], w, bf[:, :, :1].expand(-1, -1, n_pol) if shared else bf, T["freq_chan"], gpw, kernel_side_layout=True)
            iw = f()
            torch.cuda.synchronize()
            res[(tag, "iw", shared)] = iw.clone()
            out["%s iw_degrid shared=%d ms" % (tag, shared)] = round(timeit(f)[0], 4)
    for fpo in ((True, False) if n_pol == 2 else (False,)):
        a, b = res[("old", "rho", fpo)], res[("new", "rho", fpo)]
        out["rho rel diff fpo=%d" % fpo] = float((a[0] - b[0]).abs().max() / a[0].abs().max())
        out["sw rel diff fpo=%d" % fpo] = float(((a[1] - b[1]).abs() / a[1].abs().clamp_min(1e-300)).max())
        out["rho mask equal fpo=%d" % fpo] = bool(((a[0] != 0) == (b[0] != 0)).all())
    for shared in ((False, True) if n_pol == 2 else (False,)):
        a, b = res[("old", "iw", shared)], res[("new", "iw", shared)]
        # the two generations were fed their own density (1e-16 apart): compare new-kernel output on ONE density too
        out["iw max rel diff shared=%d" % shared] = float(((a - b).abs() / a.abs().clamp_min(1e-30)).nan_to_num(0).max())
        out["iw nan mask equal shared=%d" % shared] = bool((a.isnan() == b.isnan()).all())
    # bit-exactness of the degrid generations on the SAME density
    rho, rsw = res[("new", "rho", False)]
    bf = _imaging_weight.calculate_briggs_parms(rho, rsw, dict(weighting="briggs", robust=0.5))
    outs = []
    for env in ("1", "0"):
        os.environ["CNGI_IW_DEGRID_MLP"] = env
        outs.append(_imaging_weight._standard_imaging_weight_degrid_numpy_wrap(rho, T["uvw"], w, bf, T["freq_chan"], gpw,
                                                                               kernel_side_layout=True).clone())
    torch.cuda.synchronize()
    a, b = outs
    out["iw bit exact on one density"] = bool(((a == b) | (a.isnan() & b.isnan())).all())
    return out


if __name__ == "__main__":
    cases = [("f32", "continuum", 2, 500), ("f64", "continuum", 2, 120), ("f32", "cube", 2, 60), ("f32", "continuum", 1, 120),
             ("f64", "cube", 1, 40)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:1]
    allr = {}
    for c in cases:
        allr["%s %s np=%d nt=%d" % c] = run(*c)
        print(json.dumps({"%s %s np=%d nt=%d" % c: allr["%s %s np=%d nt=%d" % c]}), flush=True)
