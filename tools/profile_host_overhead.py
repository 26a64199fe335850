"""Host-side (Python) enqueue cost of one bench step, without waiting for the GPU (development tool)."""
import cProfile, pstats, io, os, sys, time
from types import SimpleNamespace
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, distributed as D
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D
d = synth.config_c2(n_time=100, dtype="f32", shard=0)
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
K = 200
t = time.perf_counter()
for _ in range(K): pipe.step(T)
t_enq = time.perf_counter() - t
pipe.flush(); torch.cuda.synchronize()
t_all = time.perf_counter() - t
print("enqueue %.3f ms/step, total %.3f ms/step" % (t_enq / K * 1e3, t_all / K * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): pipe.step(T)
pr.disable(); pipe.flush(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
