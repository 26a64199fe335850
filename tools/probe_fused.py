"""Fused image + psf pass vs the two single passes on config C2 (fp32 / fp64, continuum / 32-channel cube)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_grid as sg  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402

cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
out = {}
for prec in ("f32", "f64"):
    for mode, n_chan in (("continuum", 128), ("cube", 32)):
        d = synth.config_c2(n_time=500, dtype=prec)
        T = {k: torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan")}
        if n_chan < 128:
            T["vis"], T["weight"], T["freq_chan"] = T["vis"][:, :, :n_chan].contiguous(), T["weight"][:, :, :n_chan].contiguous(), T["freq_chan"][:n_chan].contiguous()
        gp = synth.grid_parms_for(4096, d["cell"], chan_mode=mode)
        n_ic = n_chan if mode == "cube" else 1
        rdt, cdt = (torch.float32, torch.complex64) if prec == "f32" else (torch.float64, torch.complex128)
        g = torch.zeros((n_ic, 2, 4096, 4096), dtype=cdt, device="cuda")
        pg = torch.zeros((n_ic, 2, 4096, 4096), dtype=rdt, device="cuda")
        sw = torch.zeros((n_ic, 2), dtype=torch.float64, device="cuda")
        psw = torch.zeros((n_ic, 2), dtype=torch.float64, device="cuda")
        t_img = timeit(lambda: sg.standard_grid(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, False, True, grid=g, sum_weight=sw))[0]
        t_psf = timeit(lambda: sg.standard_grid(None, T["uvw"], T["weight"], T["freq_chan"], cgk, gp, True, False, grid=pg, sum_weight=psw))[0]
        t_fused = timeit(lambda: sg.standard_grid_image_psf(T["vis"], T["uvw"], T["weight"], T["freq_chan"], cgk, gp, grid=g, sum_weight=sw, psf_grid=pg, psf_sum_weight=psw, force_fused=True))[0]
        out["%s_%s" % (prec, mode)] = {"image_ms": t_img, "psf_ms": t_psf, "fused_ms": t_fused, "samples": int(T["weight"].numel()),
                                       "saving": 1 - t_fused / (t_img + t_psf)}
        del T, g, pg, d
        torch.cuda.empty_cache()
print(json.dumps(out))
