"""The reference's own numba loops, scheduled the way its dask graph schedules them -- BENCH INFRASTRUCTURE.

bench.py's CPU legs time this (cpu_baseline.kind = "reference"): the unmodified
/root/reference/ngcasa/imaging/_imaging_utils/_standard_grid.py functions (staged under oracle/_ref on the GPU box
by oracle/make_ref.sh, loaded by oracle/ref_loader.py) called once per time chunk from a thread pool -- they are
`@jit(nopython=True, nogil=True)` (:242,:466), which is how dask's threaded scheduler runs them in parallel -- with
the partial grids summed pairwise like `_tree_sum_list` (:109-120).  Chunk graph of the C2 step:

    make_imaging_weight.py:144-247   density grid per chunk (do_imaging_weight, support 1) -> tree sum
                                     -> calculate_briggs_parms (:198-213, numpy) -> weight degrid per chunk
    make_grid.py / make_image.py     standard gridding of vis * imaging weight per chunk -> tree sum

Nothing in cngi_prototype_b200/ imports this module.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import ref_loader


def available():
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return ref_loader.available()


def _tree_sum(parts, pool=None):
    """Pairwise sum in the order of _tree_sum_list (_standard_grid.py:109-120); the adds of one level are independent
    dask tasks, so they run on the pool's threads at once (numpy releases the GIL in `+`)."""
    parts = list(parts)
    while len(parts) > 1:
        pairs = [(parts[i], parts[i + 1]) for i in range(0, len(parts) - 1, 2)]
        add = (lambda ab: ab[0] + ab[1])
        nxt = list(pool.map(add, pairs)) if pool is not None and parts[0].size > 1 << 16 else [add(ab) for ab in pairs]
        if len(parts) % 2:
            nxt.append(parts[-1])
        parts = nxt
    return parts[0]


def _chunks(n_time, n_chunks):
    n_chunks = max(1, min(n_chunks, n_time))
    b = [(n_time * i) // n_chunks for i in range(n_chunks + 1)]
    return [slice(b[i], b[i + 1]) for i in range(n_chunks) if b[i + 1] > b[i]]


class ReferenceStep:
    """make_imaging_weight (Briggs) + standard gridding of one dataset with the reference's numba functions."""

    def __init__(self, n_threads=None):
        self.sg, _, self.ck = ref_loader.load()
        self.n_threads = int(n_threads or os.cpu_count() or 1)
        self.pool = ThreadPoolExecutor(self.n_threads)

    def close(self):
        self.pool.shutdown()

    def cgk_1D(self, oversampling, support):
        return self.ck._create_prolate_spheroidal_kernel_1D(oversampling, support)

    def step(self, d, gp, gp_iw, iw_parms, cgk_1D):
        sg, pool = self.sg, self.pool
        sls = _chunks(d["weight"].shape[0], self.n_threads)
        one = np.ones(1)
        rho_parts = list(pool.map(lambda sl: sg._standard_grid_psf_numpy_wrap(d["uvw"][sl], d["weight"][sl], d["freq_chan"],
                                                                              one, gp_iw), sls))
        rho = _tree_sum([p[0] for p in rho_parts], pool)
        sw = _tree_sum([p[1] for p in rho_parts])
        del rho_parts
        robust = iw_parms["robust"]                               # calculate_briggs_parms, make_imaging_weight.py:198-213
        bf = np.ones((2,) + sw.shape)
        bf[0] = np.square(5.0 * 10.0 ** (-robust)) / (np.sum(rho ** 2, axis=(2, 3)) / sw)
        rho_api = np.moveaxis(rho, (0, 1), (2, 3))                # the API-side layout the degrid wrapper expects (:443-460)
        iw_parts = list(pool.map(lambda sl: sg._standard_imaging_weight_degrid_numpy_wrap(
            rho_api, d["uvw"][sl], d["weight"][sl], bf, d["freq_chan"], gp_iw), sls))
        g_parts = list(pool.map(lambda a: sg._standard_grid_numpy_wrap(d["vis"][a[0]], d["uvw"][a[0]], a[1], d["freq_chan"],
                                                                       cgk_1D, gp), zip(sls, iw_parts)))
        grid = _tree_sum([p[0] for p in g_parts], pool)
        sum_weight = _tree_sum([p[1] for p in g_parts])
        return grid, sum_weight
