"""N > 1 path on CPU: world_size-2 and -3 gloo jobs run the same sharding + collective control flow as the multi-GPU
bench (cngi_prototype_b200/distributed.py), with the oracle as compute, and must reproduce the single-process result."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))


def test_shard_range_partitions_exactly():
    from cngi_prototype_b200.distributed import shard_range
    for n in (0, 1, 7, 500, 1024):
        for ws in (1, 2, 3, 8):
            blocks = [shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world_size", [2, 3])
def test_sharded_step_matches_single_process(tmp_path, oracle, world_size):
    sys.path.insert(0, HERE)
    import _gloo_worker as W
    port = 29600 + world_size + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world_size),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "_gloo_worker.py"), str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:]
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world_size)]

    # single-process reference chain
    d, gp, gp_iw, n = W.dataset()
    cgk = oracle._create_prolate_spheroidal_kernel_1D(100, 7)
    rho, sw = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], np.ones(1), gp_iw)
    bf = oracle._calculate_briggs_parms(rho, sw, W.IW)
    iw = oracle._standard_imaging_weight_degrid_numpy_wrap(np.moveaxis(rho, (0, 1), (2, 3)), d["uvw"], d["weight"], bf,
                                                           d["freq_chan"], gp_iw)
    g, s = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], iw, d["freq_chan"], cgk, gp)

    # every rank saw the same all-reduced density; rank 0 holds the reduced grid
    for rk in ranks:
        assert np.max(np.abs(rk["density"] - rho)) <= 1e-12 * np.max(np.abs(rho))
    assert np.array_equal(ranks[0]["grid"] != 0, g != 0)
    assert np.max(np.abs(ranks[0]["grid"] - g)) <= 1e-12 * np.max(np.abs(g))
    assert np.max(np.abs(ranks[0]["gsw"] - s)) <= 1e-12 * np.max(np.abs(s))
    # pipelined driver: same grid on rank 0 after three overlapped steps, same imaging weights
    assert np.max(np.abs(ranks[0]["pipe_grid"] - g)) <= 1e-12 * np.max(np.abs(g))
    assert np.max(np.abs(ranks[0]["pipe_gsw"] - s)) <= 1e-12 * np.max(np.abs(s))
    for rk in ranks:   # (summation order inside the all-reduce may differ between the two drivers: tolerance, not bits)
        a, b = rk["iw_pipe"], rk["iw"]
        assert np.array_equal(np.isnan(a), np.isnan(b))
        fin = np.isfinite(b)
        assert np.max(np.abs(a[fin] - b[fin])) <= 1e-12 * np.max(np.abs(b[fin]))
    # imaging weights of the shards concatenate to the full set
    iw_cat = np.concatenate([rk["iw"] for rk in ranks], axis=0)
    assert iw_cat.shape == iw.shape
    m = np.isfinite(iw)
    assert np.array_equal(np.isnan(iw_cat), np.isnan(iw))
    assert np.max(np.abs(iw_cat[m] - iw[m])) <= 1e-12 * np.max(np.abs(iw[m]))
    # cube: channel blocks owned end to end, concatenation along the image-channel axis == unsharded cube grid
    gpc = dict(gp, chan_mode="cube")
    gc, sc = oracle._standard_grid_numpy_wrap(d["vis"], d["uvw"], d["weight"], d["freq_chan"], cgk, gpc)
    assert np.array_equal(np.concatenate([rk["cube_grid"] for rk in ranks], axis=0), gc)
    assert np.array_equal(np.concatenate([rk["cube_sw"] for rk in ranks], axis=0), sc)
    # cube imaging driver (chunked, grid -> image per chunk): channel-sharded planes concatenate to the single-process
    # cube image; the time-split variant (partial grids reduced onto the root) gives the same cube on rank 0
    gpi = dict(gpc, image_size=np.array([80, 80]))
    corr = oracle._remove_padding(oracle._create_prolate_spheroidal_image_2D(gpi["image_size_padded"]), gpi["image_size"])
    ref_img = oracle.correct_image(oracle.grid_to_uncorrected_image(gc, gpi["image_size"]), sc, corr)
    cat = np.concatenate([rk["cube_img"] for rk in ranks], axis=2)
    assert cat.shape == ref_img.shape
    assert np.max(np.abs(cat - ref_img)) <= 1e-12 * np.max(np.abs(ref_img))
    assert np.array_equal(np.concatenate([rk["cube_img_sw"] for rk in ranks], axis=0), sc)
    assert "cube_img_ts" in ranks[0] and all("cube_img_ts" not in rk for rk in ranks[1:])
    assert np.max(np.abs(ranks[0]["cube_img_ts"] - ref_img)) <= 1e-12 * np.max(np.abs(ref_img))
    assert np.max(np.abs(ranks[0]["cube_img_ts_sw"] - sc)) <= 1e-12 * np.max(np.abs(sc))
    # with_psf: the psf cube of the same samples went through the same time-split reduce
    gp_psf = dict(gpc, do_psf=True, complex_grid=False)
    pg, ps = oracle._standard_grid_psf_numpy_wrap(d["uvw"], d["weight"], d["freq_chan"], cgk, gp_psf)
    ref_psf = oracle.correct_image(oracle.grid_to_uncorrected_image(pg, gpi["image_size"]), ps, corr)
    assert np.max(np.abs(ranks[0]["cube_psf_ts"] - ref_psf)) <= 1e-12 * np.max(np.abs(ref_psf))
    assert np.max(np.abs(ranks[0]["cube_psf_ts_sw"] - ps)) <= 1e-12 * np.max(np.abs(ps))
    # rotating roots: every rank owns the chunks j with j mod world_size == rank; together they make the cube
    seen = np.zeros(6, dtype=int)
    for r, rk in enumerate(ranks):
        owned = rk["cube_rot_owned"]
        assert [int(c0) % world_size for c0, _ in owned] == [r] * len(owned) and len(owned) >= 6 // world_size
        for c0, c1 in owned:
            seen[c0:c1] += 1
            a, b = rk["cube_rot_img"][:, :, c0:c1], ref_img[:, :, c0:c1]
            assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(ref_img))
            assert np.max(np.abs(rk["cube_rot_sw"][c0:c1] - sc[c0:c1])) <= 1e-12 * np.max(np.abs(sc))
    assert np.array_equal(seen, np.ones(6, dtype=int))
