"""Device timing of the weight degrid (A4), the gridder (A1) and the fused pass (cngi_b200_standard_grid_weighted) on C2."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_grid as sg, _imaging_weight as iw  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D  # noqa: E402
from probe_std_grid import timeit  # noqa: E402


def main():
    d = synth.config_c2(dtype="f32")
    n = 4096
    out = {}
    for prec in sys.argv[1:] or ["f32"]:
        cdt, rdt = (torch.complex64, torch.float32) if prec == "f32" else (torch.complex128, torch.float64)
        vis, w = torch.as_tensor(d["vis"]).to(cdt).cuda(), torch.as_tensor(d["weight"]).to(rdt).cuda()
        uvw, freq = torch.as_tensor(d["uvw"]).cuda(), torch.as_tensor(d["freq_chan"]).cuda()
        cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(100, 7)).cuda()
        gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
        gpw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                                   complex_grid=False, do_imaging_weight=True)
        rho, sw = iw.imaging_weight_grid(uvw, w, freq, gpw, first_pol_only=True)
        bf1 = iw.calculate_briggs_parms(rho[:, :1], sw[:, :1], {"weighting": "briggs", "robust": 0.5})
        rho_x, bf_x = rho[:, :1].expand(-1, 2, -1, -1), bf1.expand(-1, -1, 2)
        grid = torch.zeros((1, 2, n, n), dtype=cdt, device="cuda")
        gsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
        iwt = iw._standard_imaging_weight_degrid_numpy_wrap(rho_x, uvw, w, bf_x, freq, gpw, kernel_side_layout=True)
        r = {}
        r["A2_density_first_pol"] = timeit(lambda: iw.imaging_weight_grid(uvw, w, freq, gpw, grid=rho, sum_weight=sw, first_pol_only=True))[0]
        r["A4_degrid"] = timeit(lambda: iw._standard_imaging_weight_degrid_numpy_wrap(rho_x, uvw, w, bf_x, freq, gpw, kernel_side_layout=True))[0]
        r["A1_grid"] = timeit(lambda: sg.standard_grid(vis, uvw, iwt, freq, cgk, gp, False, True, grid=grid, sum_weight=gsw))[0]
        for name, shared in (("fused_pol_shared", True), ("fused_per_pol", False)):
            src = dict(density=rho_x, briggs_factors=bf_x, grid_parms=gpw, pol_shared=shared)
            r[name] = timeit(lambda: sg.standard_grid(vis, uvw, w, freq, cgk, gp, False, True, grid=grid, sum_weight=gsw,
                                                      imaging_weight_from=src))[0]
        out[prec] = {k: round(v, 3) for k, v in r.items()}
        del vis, w, grid, iwt
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
