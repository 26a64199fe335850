"""Worker for tests/test_distributed_gloo.py: one rank of a world_size-N gloo job on CPU.
Runs the SAME control flow the GPU bench uses (cngi_prototype_b200.distributed.continuum_imaging_step) with the
CPU oracle injected as the compute operators."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from cngi_prototype_b200 import synth, distributed as D, read_vis as rv  # noqa: E402
from oracle import oracle as O  # noqa: E402

IW = dict(weighting="briggs", robust=0.5)


def oracle_ops():
    def iw_grid(uvw, w, freq, gp_iw, grid=None, sum_weight=None, first_pol_only=False):
        g, s = O._standard_grid_psf_numpy_wrap(uvw.numpy(), w.numpy(), freq.numpy(), np.ones(1), gp_iw)
        if first_pol_only:   # the CUDA operator updates plane 0 only; mimic that so the replicate step is exercised
            g[:, 1:] = 0
            s[:, 1:] = 0
        grid += torch.from_numpy(g)
        sum_weight += torch.from_numpy(s)

    def briggs(density, sw, parms):
        return torch.from_numpy(O._calculate_briggs_parms(density.numpy(), sw.numpy(), parms))

    def degrid(density, uvw, w, bf, freq, gp_iw):
        api = np.moveaxis(density.numpy(), (0, 1), (2, 3))
        return torch.from_numpy(O._standard_imaging_weight_degrid_numpy_wrap(api, uvw.numpy(), w.numpy(), bf.numpy(),
                                                                             freq.numpy(), gp_iw))

    def std_grid(vis, uvw, w, freq, cgk, gp, grid=None, sum_weight=None):
        g, s = O._standard_grid_numpy_wrap(vis.numpy(), uvw.numpy(), w.numpy(), freq.numpy(), cgk, gp)
        grid += torch.from_numpy(g)
        sum_weight += torch.from_numpy(s)

    def std_grid_psf(uvw, w, freq, cgk, gp, grid=None, sum_weight=None):
        g, s = O._standard_grid_psf_numpy_wrap(uvw.numpy(), w.numpy(), freq.numpy(), cgk, dict(gp, do_psf=True, complex_grid=False))
        grid += torch.from_numpy(g)
        sum_weight += torch.from_numpy(s)

    def zeros(shape, is_complex):
        return torch.zeros(shape, dtype=torch.complex128 if is_complex else torch.float64)

    def to_image(g, s, gp):
        corr = O._remove_padding(O._create_prolate_spheroidal_image_2D(gp["image_size_padded"]), gp["image_size"])
        img = O.correct_image(O.grid_to_uncorrected_image(g.numpy(), gp["image_size"]), s.numpy(), corr)
        return torch.from_numpy(np.ascontiguousarray(img))

    return SimpleNamespace(imaging_weight_grid=iw_grid, briggs=briggs, degrid=degrid, standard_grid=std_grid, standard_grid_psf=std_grid_psf,
                           zeros=zeros, to_image=to_image)


def dataset():
    d = synth.make_vis_set(7, 24, 6, 2, 1e9, 1.1e9, 300.0, 200.0, seed=17)
    n = 96
    gp = synth.grid_parms_for(n, d["cell"], chan_mode="continuum")
    gp_iw = synth.grid_parms_for(n, d["cell"], chan_mode="continuum", support=1, oversampling=0, do_psf=True,
                                 complex_grid=False, do_imaging_weight=True)
    return d, gp, gp_iw, n


def main():
    out = sys.argv[1]
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    d, gp, gp_iw, n = dataset()
    cgk = O._create_prolate_spheroidal_kernel_1D(100, 7)
    full = {k: torch.from_numpy(np.ascontiguousarray(d[k])) for k in ("vis", "uvw", "weight", "freq_chan")}
    # continuum: time sharding + all-reduce(density) + reduce(grid)
    # ... read by every rank from a zarr store: lazy arrays, so a rank decodes only the chunk files of its time shard
    store = os.path.join(out, "sim.vis.zarr")
    if rank == 0:
        rv.write_vis(store, {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"], "chan": d["freq_chan"]},
                     chunks={"time": 5, "chan": 4})
    dist.barrier()
    xds = rv.read_vis(store, partition="xds0").xds0
    lazy = {"vis": xds["DATA"], "uvw": xds["UVW"], "weight": xds["WEIGHT"], "freq_chan": xds["chan"]}
    shard = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in D.time_shard(lazy, rank, ws).items()}
    for k, v in D.time_shard(full, rank, ws).items():
        assert torch.equal(torch.nan_to_num(shard[k].view(torch.float64) if shard[k].is_complex() else shard[k]),
                           torch.nan_to_num(v.view(torch.float64) if v.is_complex() else v)), k
    bufs = SimpleNamespace(density=torch.zeros((1, 2, n, n), dtype=torch.float64), dsw=torch.zeros((1, 2), dtype=torch.float64),
                           grid=torch.zeros((1, 2, n, n), dtype=torch.complex128), gsw=torch.zeros((1, 2), dtype=torch.float64))
    iw = D.continuum_imaging_step(oracle_ops(), shard, gp, gp_iw, IW, cgk, bufs)
    # the software-pipelined driver (collectives overlapped across steps) must give the same result every step
    def make_bufs():
        return SimpleNamespace(density=torch.zeros((1, 2, n, n), dtype=torch.float64), dsw=torch.zeros((1, 2), dtype=torch.float64),
                               grid=torch.zeros((1, 2, n, n), dtype=torch.complex128), gsw=torch.zeros((1, 2), dtype=torch.float64))
    pipe = D.ContinuumPipeline(oracle_ops(), gp, gp_iw, IW, cgk, make_bufs)
    for _ in range(3):
        pipe.step(shard)
    iw_pipe = pipe.flush()
    pipe_grid, pipe_gsw = pipe.last.grid.numpy().copy(), pipe.last.gsw.numpy().copy()
    # cube: channel sharding, no exchange; gather the owned planes only to check them
    gpc = dict(gp, chan_mode="cube")
    cs = D.channel_shard(full, rank, ws)
    gc, sc = O._standard_grid_numpy_wrap(cs["vis"].numpy(), cs["uvw"].numpy(), cs["weight"].numpy(), cs["freq_chan"].numpy(),
                                         cgk, gpc)
    # cube imaging driver: channel groups own their planes end to end, chunked through one grid buffer; with
    # time_split == world size the ranks of the single group reduce their partial chunk grids onto the root
    gpi = dict(gpc, image_size=np.array([80, 80]))
    img1, sw1, cr1 = D.cube_imaging(oracle_ops(), full, gpi, cgk, chan_chunk=2, time_split=1)
    groups = D.make_time_groups(ws, ws)
    img2, sw2, psf2, psw2, cr2 = D.cube_imaging(oracle_ops(), full, gpi, cgk, chan_chunk=4, time_split=ws, groups=groups,
                                                with_psf=True)
    # rotating roots: chunk j of the single group is reduced onto rank j mod ws, every rank transforms the chunks it owns
    img3, sw3, own3 = D.cube_imaging(oracle_ops(), full, gpi, cgk, chan_chunk=1, time_split=ws, groups=groups, rotate_roots=True)
    extra = dict(cube_img=img1.numpy(), cube_img_sw=sw1.numpy(), cube_img_range=np.array(cr1),
                 cube_rot_owned=np.array(own3).reshape(-1, 2))
    if img3 is not None:
        extra.update(cube_rot_img=img3.numpy(), cube_rot_sw=sw3.numpy())
    if img2 is not None:
        extra.update(cube_img_ts=img2.numpy(), cube_img_ts_sw=sw2.numpy(), cube_psf_ts=psf2.numpy(), cube_psf_ts_sw=psw2.numpy())
    np.savez(os.path.join(out, "rank%d.npz" % rank), **extra, grid=bufs.grid.numpy(), gsw=bufs.gsw.numpy(), iw=iw.numpy(),
             density=bufs.density.numpy(), pipe_grid=pipe_grid, pipe_gsw=pipe_gsw, iw_pipe=iw_pipe.numpy(), cube_grid=gc, cube_sw=sc, chan_range=np.array(D.shard_range(6, rank, ws)),
             time_range=np.array(D.shard_range(24, rank, ws)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
