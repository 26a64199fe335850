"""Quick device-timing probe of the aperture gridder on config C3 (development tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _aperture_grid  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402

d = synth.config_c2(n_time=200, n_chan=64, dtype="f32")
gcf = synth.make_mosaic_gcf(d["n_baseline"], 64, 2, n_field=7)
d["field"] = synth.mosaic_field_column(d["uvw"].shape[0], d["n_baseline"], gcf["field_id"])
gp = synth.grid_parms_for(2048, d["cell"] * 1.1, chan_mode="continuum")
gp["oversampling"], gp["field_id"] = gcf["oversampling"], gcf["field_id"]
T = {k: torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan", "field")}
G = {k: torch.as_tensor(v).cuda() for k, v in gcf.items()}
grid = torch.zeros((1, 2, 2048, 2048), dtype=torch.complex64, device="cuda")
gsw = torch.zeros((1, 2), dtype=torch.float64, device="cuda")
common = (T["uvw"], T["weight"], T["field"], G["cf_baseline_map"], G["cf_chan_map"], G["cf_pol_map"])
print("samples %.1fM" % (d["weight"].size / 1e6))
print("aperture image", timeit(lambda: _aperture_grid._aperture_grid_numpy_wrap(
    T["vis"], *common, G["conv_kernel"], gcf["weight_support"], G["phase_gradient"], T["freq_chan"], gp, grid=grid,
    sum_weight=gsw)))
print("aperture weight", timeit(lambda: _aperture_grid._aperture_weight_grid_numpy_wrap(
    *common, G["weight_conv_kernel"], gcf["weight_support"], G["phase_gradient"], T["freq_chan"], gp, grid=grid,
    sum_weight=gsw)))
