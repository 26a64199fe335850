"""Generates the fixtures of the SURVEY.md section 8(f) rows by running the UNMODIFIED reference (this container only).

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/make_golden_next.py [direction_rotate] [gcf] [pb]

direction_rotate_*.npz : /root/reference/ngcasa/imaging/direction_rotate.py:127-248
                         (calc_rotation_mats, apply_rotation_matrix, apply_phasor; the xarray FIELD table is
                         replaced by a 10-line stand-in that implements the one `.sel(field_id=, d1=0)` call)
gcf_*.npz              : /root/reference/ngcasa/imaging/make_gridding_convolution_function.py:331-457,512-560
                         and _imaging_utils/_make_pb_symmetric.py (see the gcf section below)
pb_*.npz               : _imaging_utils/_make_pb_symmetric.py:26-132 (_airy_disk, _casa_airy_disk), ipower 1 and 2
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402

REF = ref_loader.REF_ROOT


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-32s %8.1f KB" % (name, os.path.getsize(path) / 1024))


class _Values:
    def __init__(self, a):
        self.values = a

    def __getitem__(self, k):
        return _Values(self.values[k])


class _PhaseDir:
    """Stand-in for the xarray FIELD.PHASE_DIR variable: only `.sel(field_id=ids, d1=0)` is used
    (direction_rotate.py:153)."""

    def __init__(self, ids, dirs):
        self.ids, self.dirs = list(ids), dirs

    def sel(self, field_id, d1):
        return _Values(self.dirs[[self.ids.index(f) for f in field_id]])


class _NS:
    pass


def direction_rotate():
    ref_loader.load()
    dr = ref_loader._load("ref_direction_rotate", os.path.join(REF, "ngcasa", "imaging", "direction_rotate.py"))
    rng = np.random.default_rng(31)
    n_t, n_b, n_c, n_p = 9, 7, 5, 2
    table_ids = np.array([2, 3, 7, 9])                      # FIELD table has one field the data never uses
    table_dirs = np.array([[0.2, -0.1], [1.0, 0.5], [1.0011, 0.50052], [0.99893, 0.49961]])
    field_of_time = np.array([3, 7, 9])[np.arange(n_t) % 3]
    field = np.repeat(field_of_time[:, None], n_b, 1).astype(np.int64)
    field[1, 2] = -2147483648                               # INT_NAN rows are ignored by both lookups
    field[4, 0] = -2147483648
    uvw = rng.normal(0.0, 300.0, (n_t, n_b, 3))
    uvw[:, :, 2] *= 0.1
    vis = rng.standard_normal((n_t, n_b, n_c, n_p)) + 1j * rng.standard_normal((n_t, n_b, n_c, n_p))
    vis[0, 1, 2, 0] = np.nan
    freq = np.linspace(345.0e9, 345.4e9, n_c)
    fd, vd = _NS(), _NS()
    fd.PHASE_DIR = _PhaseDir(table_ids, table_dirs)
    vd.FIELD_ID = field
    for tag, ctr, sp in (("ctr_sp", True, True), ("full_dp", False, False)):
        rp = {"new_phase_center": np.array([1.0003, 0.5002]), "common_tangent_reprojection": ctr,
              "single_precision": sp}
        rotmat, phase_rot, ids = dr.calc_rotation_mats(vd, fd, rp)
        uvw_rot = dr.apply_rotation_matrix(uvw, field[:, :, None], rotmat, ids)
        vis_rot = dr.apply_phasor(vis, uvw_rot[:, :, :, None], field[:, :, None, None], freq[None, None, :, None],
                                  phase_rot, ids, ctr, sp)
        save("direction_rotate_" + tag, uvw=uvw, vis=vis, field=field, freq_chan=freq, table_ids=table_ids,
             table_dirs=table_dirs, new_phase_center=rp["new_phase_center"], ctr=np.array(ctr), sp=np.array(sp),
             uvw_rotmat=rotmat, phase_rotation=phase_rot, rot_field_id=ids, uvw_rot=uvw_rot, vis_rot=vis_rot)


def gcf():
    """Runs the reference's own helpers on one small heterogeneous-array case (2 dish types -> 3 antenna pairs,
    2 PB frequencies): create_cf_baseline_map :512, create_cf_chan_map :536, make_baseline_patterns :394 with
    _casa_airy_disk_rorder / _airy_disk_rorder (_make_pb_symmetric.py:187,135), the fft of :246-247 with numpy.fft
    (dask.array.fft wraps it), resize_and_calc_support :361.  make_phase_gradient :331 needs astropy (absent):
    not in the fixture."""
    ref_loader.load()
    mg = ref_loader._load("ref_make_gcf", os.path.join(REF, "ngcasa", "imaging",
                                                       "make_gridding_convolution_function.py"))
    pbm = ref_loader._load("ref_make_pb_symmetric", os.path.join(REF, "ngcasa", "imaging", "_imaging_utils",
                                                                 "_make_pb_symmetric.py"))
    rng = np.random.default_rng(41)
    n_ant = 9
    unique_ant_indx = np.array([0, 0, 0, 1, 0, 1, 0, 0, 1])
    a1, a2 = np.triu_indices(n_ant, 1)
    baseline_ant = np.stack([a1, a2], 1)
    freq_chan = np.linspace(100.0e9, 101.1e9, 12)
    for tag, func, n_pad, cell_as in (("casa_airy", pbm._casa_airy_disk_rorder, np.array([240, 256]), 0.55),
                                      ("airy", pbm._airy_disk_rorder, np.array([250, 250]), 0.5)):
        cell = np.array([-cell_as, cell_as]) * np.pi / (180 * 3600)
        gcf_parms = dict(list_dish_diameters=np.array([10.7, 6.25]), list_blockage_diameters=np.array([0.75, 0.0 if tag == "airy" else 0.75]),
                         unique_ant_indx=unique_ant_indx, basline_ant=baseline_ant, freq_chan=freq_chan,
                         pol=np.array([0, 1]), oversampling=np.array([5, 5]), max_support=np.array([11, 11]),
                         support_cut_level=0.025, chan_tolerance_factor=0.005)
        gcf_parms["resize_conv_size"] = (gcf_parms["max_support"] + 1) * gcf_parms["oversampling"]
        grid_parms = dict(image_size=n_pad, image_size_padded=n_pad, cell_size=cell, image_center=n_pad // 2)
        cf_bl_map, pairs = mg.create_cf_baseline_map(unique_ant_indx, baseline_ant, 2)
        cf_chan_map, pb_freq = mg.create_cf_chan_map(freq_chan, 0.005)
        planes = {}
        for ipower in (1, 2):
            gcf_parms["ipower"] = ipower
            bp = mg.make_baseline_patterns(pb_freq, np.array([0]), pairs, func, gcf_parms, grid_parms)
            planes[ipower] = np.real(np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(bp, axes=(3, 4)), axes=(3, 4)),
                                                     axes=(3, 4)))
        ck, wk, sup = mg.resize_and_calc_support(planes[1], planes[2], gcf_parms, grid_parms)
        save("gcf_" + tag, n_pad=n_pad, cell_size=cell, unique_ant_indx=unique_ant_indx, baseline_ant=baseline_ant,
             freq_chan=freq_chan, dish=gcf_parms["list_dish_diameters"], blockage=gcf_parms["list_blockage_diameters"],
             oversampling=gcf_parms["oversampling"], max_support=gcf_parms["max_support"],
             cf_baseline_map=cf_bl_map, pb_ant_pairs=pairs, cf_chan_map=cf_chan_map, pb_freq=pb_freq,
             conv_kernel=ck, weight_conv_kernel=wk, support=sup,
             centre_row_pb=planes[1][:, :, 0, :, n_pad[1] // 2])
    # create_cf_chan_map corner cases (:536-560)
    cases = {}
    for k, (f, tol) in enumerate(((np.linspace(1e9, 2e9, 64), 0.005), (np.linspace(1e9, 1.001e9, 8), 0.005),
                                  (np.array([1.4e9]), 0.005), (np.linspace(1e9, 2e9, 7), 0.2))):
        m, pf = mg.create_cf_chan_map(f, tol)
        cases["f%d" % k], cases["tol%d" % k], cases["map%d" % k], cases["pbf%d" % k] = f, np.array(tol), m, pf
    save("gcf_chan_maps", **cases)


def pb():
    """_airy_disk / _casa_airy_disk (_make_pb_symmetric.py:26,79) as make_pb.py:95-107 calls them (ipower 2)."""
    ref_loader.load()
    pbm = ref_loader._load("ref_make_pb_symmetric", os.path.join(REF, "ngcasa", "imaging", "_imaging_utils",
                                                                 "_make_pb_symmetric.py"))
    freq = np.array([100.0e9, 100.7e9, 101.5e9])
    pol = np.array([0, 1])
    grid_parms = dict(image_size=np.array([24, 21]), image_center=np.array([12, 10]),
                      cell_size=np.array([-3.1, 3.1]) * np.pi / (180 * 3600))
    for tag, func, block in (("casa_airy", pbm._casa_airy_disk, [0.75, 0.0]), ("airy", pbm._airy_disk, [0.75, 0.5])):
        pb_parms = dict(list_dish_diameters=[10.7, 6.25], list_blockage_diameters=block, ipower=2)
        out = func(freq, pol, pb_parms, grid_parms)
        pb_parms["ipower"] = 1
        out1 = func(freq, pol, pb_parms, grid_parms)
        save("pb_" + tag, freq_chan=freq, pol=pol, image_size=grid_parms["image_size"], image_center=grid_parms["image_center"],
             cell_size=grid_parms["cell_size"], dish=np.array(pb_parms["list_dish_diameters"]), blockage=np.array(block),
             pb=out, voltage=out1)


if __name__ == "__main__":
    which = sys.argv[1:] or ["direction_rotate", "gcf", "pb"]
    for w in which:
        globals()[w]()
