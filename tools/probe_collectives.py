"""Under torchrun: time the two collectives of the time-sharded continuum step alone, and the pipelined step with the
grid reduce done as reduce-to-root / rotating root / all-reduce.  One JSON line from rank 0."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import synth, distributed as D  # noqa: E402
from cngi_prototype_b200._gridding_convolutional_kernels import _create_prolate_spheroidal_kernel_1D  # noqa: E402


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return round(float(ms.item()), 4)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("PROBE_PLAIN_INIT"):
        dist.init_process_group("nccl", device_id=dev)
    else:
        D.init_nccl(dev)
    n = 4096
    out = {"world": world, "NCCL_ALGO": os.environ.get("NCCL_ALGO", "default"),
           "knobs": {k: v for k, v in os.environ.items() if k.startswith(("NCCL_", "TORCH_NCCL")) and k != "NCCL_ALGO"}}
    dens = torch.zeros((1, 1, n, n), dtype=torch.float64, device=dev)
    grid = torch.zeros((1, 2, n, n, 2), dtype=torch.float32, device=dev)
    out["allreduce_density_134MB_f64_ms"] = timed(lambda: dist.all_reduce(dens))
    out["reduce_grid_268MB_f32_ms"] = timed(lambda: dist.reduce(grid, 0))
    out["allreduce_grid_268MB_f32_ms"] = timed(lambda: dist.all_reduce(grid))
    k = [0]

    def rot():
        dist.reduce(grid, k[0] % world)
        k[0] += 1
    out["reduce_grid_rotating_root_ms"] = timed(rot)
    del dens, grid
    # the pipelined step (bench.py's headline loop) with the three forms of the grid reduce
    d = synth.config_c2(n_time=500, n_chan=128, dtype="f32", shard=rank)
    T = {k2: torch.as_tensor(d[k2]).to(dev) for k2 in ("vis", "uvw", "weight", "freq_chan")}
    cgk = torch.as_tensor(_create_prolate_spheroidal_kernel_1D(100, 7)).to(dev)
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    gp_iw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)

    def make_bufs():
        return SimpleNamespace(density=torch.empty((1, 2, n, n), dtype=torch.float64, device=dev),
                               dsw=torch.empty((1, 2), dtype=torch.float64, device=dev),
                               grid=torch.empty((1, 2, n, n), dtype=torch.complex64, device=dev),
                               gsw=torch.empty((1, 2), dtype=torch.float64, device=dev))
    modes = os.environ.get("PROBE_MODES", "root0,rotate,allreduce,none").split(",")
    sym = D.SymmetricCollectives(dev, n_blocks=int(os.environ.get("PROBE_MM_BLOCKS", "0"))) if D.SymmetricCollectives.supported(dev) else None
    out["multicast_supported"] = sym is not None

    def make_bufs_sym():
        return SimpleNamespace(density=sym.empty((1, 2, n, n), torch.float64), dsw=torch.empty((1, 2), dtype=torch.float64, device=dev),
                               grid=sym.empty((1, 2, n, n), torch.complex64), gsw=torch.empty((1, 2), dtype=torch.float64, device=dev))
    if sym is not None and "multimem" in modes:
        # correctness of the two switch-side sums against NCCL, then their stand-alone times
        a = sym.empty((1, 2, n, n), torch.complex64)
        ref = torch.empty_like(a)
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        torch.view_as_real(a).normal_(generator=g)
        ref.copy_(a)
        dist.reduce(torch.view_as_real(ref), 0)
        sym.reduce_grid(a, 0).wait()
        torch.cuda.synchronize()
        if rank == 0:
            out["multimem_reduce_max_rel_diff_vs_nccl"] = float((a - ref).abs().max() / ref.abs().max())
        dd = sym.empty((1, 2, n, n), torch.float64)
        dd.normal_(generator=g)
        dref = dd[:, :1].clone()
        dist.all_reduce(dref)
        sym.allreduce_density(dd, n * n).wait()
        torch.cuda.synchronize()
        out["multimem_allreduce_max_rel_diff_vs_nccl"] = float((dd[:, :1] - dref).abs().max() / dref.abs().max())
        out["multimem_reduce_grid_268MB_ms"] = timed(lambda: sym.reduce_grid(a, 0).wait())
        out["multimem_allreduce_density_134MB_ms"] = timed(lambda: sym.allreduce_density(dd, n * n).wait())
        del a, ref, dd, dref
    for mode in modes:
        if mode == "multimem":
            if sym is None:
                continue
            pipe = D.ContinuumPipeline(D.cuda_ops(), gp, gp_iw, dict(weighting="briggs", robust=0.5), cgk, make_bufs_sym,
                                       symmetric=sym)
        else:
            pipe = D.ContinuumPipeline(D.cuda_ops(), gp, gp_iw, dict(weighting="briggs", robust=0.5), cgk, make_bufs, grid_reduce=mode)

        def run():
            for _ in range(10):
                pipe.step(T)
            pipe.flush()
        out["step_ms_grid_reduce_" + mode] = round(timed(run, n=3, warm=1) / 10, 4)
        del pipe
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
