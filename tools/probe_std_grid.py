"""Quick device-timing probe of the standard gridder kernels (development tool, not the bench)."""
import argparse
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, _standard_grid  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--n-time", type=int, default=0)
    ap.add_argument("--n-chan", type=int, default=0)
    ap.add_argument("--n-uv", type=int, default=0)
    ap.add_argument("--modes", default="continuum,cube")
    ap.add_argument("--precs", default="f32,f64")
    ap.add_argument("--algos", default="2,1")
    ap.add_argument("--groups", default="0")
    ap.add_argument("--segments", default="0")
    ap.add_argument("--psf", action="store_true")
    a = ap.parse_args()
    if a.config == "c1":
        d = synth.config_c1(n_time=a.n_time or 1000, n_chan=a.n_chan or 64, dtype="f64")
        n_uv = a.n_uv or 1024
    else:
        d = synth.config_c2(n_time=a.n_time or 500, n_chan=a.n_chan or 128, dtype="f64")
        n_uv = a.n_uv or 4096
    cgk = _create_prolate_spheroidal_kernel_1D(100, 7)
    n_samp = d["weight"].size
    print("samples %.1fM n_uv %d" % (n_samp / 1e6, n_uv), flush=True)
    for prec in a.precs.split(","):
        cdt, rdt = (torch.complex64, torch.float32) if prec == "f32" else (torch.complex128, torch.float64)
        vis = torch.as_tensor(d["vis"]).to(cdt).cuda()
        w = torch.as_tensor(d["weight"]).to(rdt).cuda()
        uvw = torch.as_tensor(d["uvw"]).cuda()
        freq = torch.as_tensor(d["freq_chan"]).cuda()
        cg = torch.as_tensor(cgk).cuda()
        for mode in a.modes.split(","):
            gp = synth.grid_parms_for(n_uv, d["cell"], chan_mode=mode)
            n_ic = vis.shape[2] if mode == "cube" else 1
            cell_b = (8 if prec == "f32" else 16)
            if n_ic * 2 * n_uv * n_uv * cell_b > 60e9:
                print("skip", prec, mode, "grid too large")
                continue
            grid = torch.zeros((n_ic, 2, n_uv, n_uv), dtype=cdt, device="cuda")
            sw = torch.zeros((n_ic, 2), dtype=torch.float64, device="cuda")
            for algo in [int(x) for x in a.algos.split(",")]:
                for G in [int(x) for x in a.groups.split(",")]:
                    for seg in [int(x) for x in a.segments.split(",")]:
                        if algo == 1 and (G != int(a.groups.split(",")[0]) or seg != int(a.segments.split(",")[0])):
                            continue
                        fn = lambda: _standard_grid.standard_grid(vis, uvw, w, freq, cg, gp, False, True, algorithm=algo,
                                                                  chan_group=G, time_segment=seg, grid=grid, sum_weight=sw)
                        med, best = timeit(fn)
                        print(json.dumps(dict(prec=prec, mode=mode, algo=algo, G=G, seg=seg, ms=round(med, 3),
                                              best_ms=round(best, 3), gvis_s=round(n_samp / med / 1e6, 2),
                                              gtap_s=round(49 * n_samp / med / 1e6, 1))), flush=True)
            del grid
        del vis, w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
