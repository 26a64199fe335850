"""CPU-only: the host-side halves of the SURVEY.md section 8(f) rows (no kernel is launched) against fixtures produced by
the reference's own functions (tests/golden/make_golden_next.py): calc_rotation_mats (direction_rotate.py:127-175),
create_cf_baseline_map / create_cf_chan_map (make_gridding_convolution_function.py:512-560), and the analytic SIN
world2pix against the CASA trace quoted in the reference (:565-577)."""
import os

import numpy as np
import pytest

from _util import GOLDEN


@pytest.mark.parametrize("name", ["direction_rotate_ctr_sp", "direction_rotate_full_dp"])
def test_calc_rotation_mats(name):
    from cngi_prototype_b200 import direction_rotate as dr
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    R, P, ids = dr.calc_rotation_mats(d["field"], d["table_ids"], d["table_dirs"],
                                      dict(new_phase_center=d["new_phase_center"],
                                           common_tangent_reprojection=bool(d["ctr"])))
    assert np.array_equal(ids, d["rot_field_id"])
    np.testing.assert_allclose(R, d["uvw_rotmat"], rtol=0, atol=1e-15)      # scipy builds them from quaternions
    np.testing.assert_allclose(P, d["phase_rotation"], rtol=0, atol=1e-18)
    if bool(d["ctr"]):
        assert np.all(R[:, 2, 0:2] == 0.0)                                  # common tangent (:166-167)


@pytest.mark.parametrize("tag", ["casa_airy", "airy"])
def test_cf_maps(tag):
    from cngi_prototype_b200 import make_gridding_convolution_function as mg
    d = np.load(os.path.join(GOLDEN, "gcf_%s.npz" % tag))
    m, pairs = mg.create_cf_baseline_map(d["unique_ant_indx"], d["baseline_ant"], len(d["dish"]))
    assert np.array_equal(m, d["cf_baseline_map"]) and np.array_equal(pairs, d["pb_ant_pairs"])
    cm, pf = mg.create_cf_chan_map(d["freq_chan"], 0.005)
    assert np.array_equal(cm, d["cf_chan_map"]) and np.array_equal(pf, d["pb_freq"])
    c = np.load(os.path.join(GOLDEN, "gcf_chan_maps.npz"))
    for k in range(4):
        cm, pf = mg.create_cf_chan_map(c["f%d" % k], float(c["tol%d" % k]))
        assert np.array_equal(cm, c["map%d" % k]) and np.array_equal(pf, c["pbf%d" % k])


def test_sin_projection_known_answer():
    from cngi_prototype_b200 import make_gridding_convolution_function as mg
    deg = np.pi / 180
    off = mg._sin_offset_in_pixels(np.array([[-179.5337374791666889 * deg, -18.863873258333338612 * deg]]),
                                   np.array([180.46846189999996568 * deg, -18.863873247222226581 * deg]),
                                   np.array([-5e-05 * deg, 5e-05 * deg]))
    np.testing.assert_allclose(off[0] + 120, [161.6249842951184803, 119.99951947142589859], rtol=0, atol=2e-9)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cngi_prototype_b200 import _lib, direction_rotate as dr, make_gridding_convolution_function as mg
    with pytest.raises(_lib.CngiError):
        dr.apply_rotation_matrix(np.zeros((1, 1, 3)), np.zeros((1, 1), dtype=np.int64), np.eye(3)[None], np.array([0]))
    with pytest.raises(_lib.CngiError):
        mg.make_phase_gradient(np.array([[1.0, 0.5]]), dict(oversampling=[5, 5], resize_conv_size=np.array([10, 10]),
                                                            phase_center=np.array([1.0, 0.5])),
                               dict(image_size_padded=np.array([64, 64]), cell_size=np.array([-1e-6, 1e-6])))
