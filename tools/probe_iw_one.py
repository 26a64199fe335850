"""One launch of each imaging-weight kernel generation on config C2 (fp32) -- the process ncu captures (development tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _imaging_weight  # noqa: E402

d = synth.config_c2(n_time=500, dtype="f32")
T = {k: torch.as_tensor(d[k]).cuda() for k in ("uvw", "weight", "freq_chan")}
gpw = synth.grid_parms_for(4096, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                           complex_grid=False, do_imaging_weight=True)
for env in ("1", "0"):
    os.environ["CNGI_IW_GRID_OLD"] = env
    os.environ["CNGI_IW_DEGRID_MLP"] = env
    rho = torch.zeros((1, 2, 4096, 4096), dtype=torch.float64, device="cuda")
    rsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
    _imaging_weight.imaging_weight_grid(T["uvw"], T["weight"], T["freq_chan"], gpw, grid=rho, sum_weight=rsw, first_pol_only=True)
    bf = _imaging_weight.calculate_briggs_parms(rho[:, :1], rsw[:, :1], dict(weighting="briggs", robust=0.5)).expand(-1, -1, 2)
    _imaging_weight._standard_imaging_weight_degrid_numpy_wrap(rho[:, :1].expand(-1, 2, -1, -1), T["uvw"], T["weight"], bf,
                                                               T["freq_chan"], gpw, kernel_side_layout=True)
    torch.cuda.synchronize()
