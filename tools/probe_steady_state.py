"""Per-chunk device timing of the bench step over a long run (is the step time stable?).  Development tool."""
import os, sys, time, subprocess
from types import SimpleNamespace
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, distributed as D
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
d = synth.config_c2(dtype="f32", shard=0)
n = 4096
gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
gp_iw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True, complex_grid=False, do_imaging_weight=True)
T = {k: torch.as_tensor(d[k]).cuda() for k in ("vis", "uvw", "weight", "freq_chan")}
cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(100, 7)).cuda()
def mk():
    return SimpleNamespace(density=torch.empty((1, 2, n, n), dtype=torch.float64, device="cuda"), dsw=torch.empty((1, 2), dtype=torch.float64, device="cuda"),
                           grid=torch.empty((1, 2, n, n), dtype=torch.complex64, device="cuda"), gsw=torch.empty((1, 2), dtype=torch.float64, device="cuda"))
pipe = D.ContinuumPipeline(D.cuda_ops(), gp, gp_iw, dict(weighting="briggs", robust=0.5), cgk, mk)
for _ in range(5): pipe.step(T)
pipe.flush(); torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(16)]
evs[0].record()
for c in range(15):
    for _ in range(20): pipe.step(T)
    evs[c + 1].record()
pipe.flush(); torch.cuda.synchronize()
print("ms/step per 20-step chunk:", [round(evs[i].elapsed_time(evs[i + 1]) / 20, 3) for i in range(15)])
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader"], stdout=subprocess.PIPE, text=True).stdout)
