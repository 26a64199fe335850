"""CUDA direction_rotate (cngi_b200_direction_rotate) against the reference fixtures and the oracle.

Reference: /root/reference/ngcasa/imaging/direction_rotate.py:127-248.  Tolerances: uvw 1e-15 relative (three
products, the reference's BLAS may fuse them); vis: the phase argument y = 2 pi d f / c is reproducible to a few
ulp(y) in the reference itself, so |delta vis| <= 8 eps |y| |vis| (+1e-12 relative, the fp64 bar of north_star).
"""
import os

import numpy as np
import pytest

from _util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: d[k] for k in d.files}


def _ybound(uvw_rot, P, freq):
    y = 2 * np.pi * np.abs(uvw_rot).sum(-1).max() * np.abs(P).max() * np.max(freq) / 299792458.0
    return 1e-12 + 8 * np.finfo(np.float64).eps * y


@pytest.mark.parametrize("name", ["direction_rotate_ctr_sp", "direction_rotate_full_dp"])
def test_golden(name):
    from cngi_prototype_b200 import direction_rotate as dr
    d = _golden(name)
    ctr, sp = bool(d["ctr"]), bool(d["sp"])
    R, P, ids = dr.calc_rotation_mats(d["field"], d["table_ids"], d["table_dirs"],
                                      dict(new_phase_center=d["new_phase_center"], common_tangent_reprojection=ctr))
    assert np.array_equal(ids, d["rot_field_id"])
    np.testing.assert_allclose(R, d["uvw_rotmat"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(P, d["phase_rotation"], rtol=0, atol=1e-18)
    # reference call sequence: rotate, then phasor on the rotated uvw
    ur = dr.apply_rotation_matrix(d["uvw"], d["field"][:, :, None], d["uvw_rotmat"], d["rot_field_id"])
    assert rel_err(ur, d["uvw_rot"]) < 1e-15
    vr = dr.apply_phasor(d["vis"], d["uvw_rot"][:, :, :, None], d["field"][:, :, None, None],
                         d["freq_chan"][None, None, :, None], d["phase_rotation"], d["rot_field_id"], ctr, sp)
    assert np.array_equal(np.isnan(vr), np.isnan(d["vis_rot"]))
    m = ~np.isnan(vr)
    bound = _ybound(d["uvw_rot"], d["phase_rotation"], d["freq_chan"])
    assert np.max(np.abs(vr[m] - d["vis_rot"][m])) <= bound * np.abs(d["vis"][m]).max()
    if sp:
        assert np.mean(vr[m] != d["vis_rot"][m]) < 1e-3      # complex64 round trip: at most a rare rounding flip
    # fused form gives the same answers
    vr2, ur2 = dr.rotate_chunk(d["vis"], d["uvw"], d["field"], d["freq_chan"], d["uvw_rotmat"], d["phase_rotation"],
                               d["rot_field_id"], ctr, sp)
    assert np.array_equal(ur2, ur)
    assert np.max(np.abs(vr2[m] - d["vis_rot"][m])) <= bound * np.abs(d["vis"][m]).max()


@pytest.mark.parametrize("n_pol,dtype,ctr", [(2, np.complex128, True), (1, np.complex128, False),
                                             (2, np.complex64, True), (4, np.complex64, False),
                                             (3, np.complex128, True)])
def test_against_oracle(oracle, n_pol, dtype, ctr):
    from cngi_prototype_b200 import direction_rotate as dr
    rng = np.random.default_rng(7 + n_pol)
    n_t, n_b, n_c = 37, 45, 19
    ids = np.array([0, 4, 5, 11, 12, 13, 20])
    dirs = np.stack([1.0 + rng.normal(0, 4e-4, 7), 0.5 + rng.normal(0, 4e-4, 7)], 1)
    field = np.repeat(ids[rng.integers(0, 7, n_t)][:, None], n_b, 1).astype(np.int64)
    field[rng.random((n_t, n_b)) < 0.05] = -2147483648
    field[:, 0] = field.max(1)                                   # keep one valid row per integration
    uvw = rng.normal(0, 300.0, (n_t, n_b, 3))
    vis = (rng.standard_normal((n_t, n_b, n_c, n_pol)) + 1j * rng.standard_normal((n_t, n_b, n_c, n_pol))).astype(dtype)
    vis[rng.random(vis.shape) < 0.02] = np.nan
    freq = np.linspace(345e9, 347e9, n_c)
    parms = dict(new_phase_center=np.array([1.0001, 0.50003]), common_tangent_reprojection=ctr)
    R, P, rid = dr.calc_rotation_mats(field, ids, dirs, parms)
    Ro, Po, rido = oracle.calc_rotation_mats(field, ids, dirs, parms["new_phase_center"], ctr)
    assert np.array_equal(R, Ro) and np.array_equal(P, Po) and np.array_equal(rid, rido)
    vr, ur = dr.rotate_chunk(vis, uvw, field, freq, R, P, rid, ctr, False)
    uo = oracle.apply_rotation_matrix(uvw, field, R, rid)
    assert np.array_equal(ur, uo)                                # same operation order, no contraction: bit-exact
    vo = oracle.apply_phasor(vis.astype(np.complex128), uo, field, freq, P, rid, ctr, False)
    assert vr.dtype == dtype
    assert np.array_equal(np.isnan(vr), np.isnan(vo))
    m = ~np.isnan(vo)
    tol = 1e-6 if dtype == np.complex64 else _ybound(uo, P, freq)
    assert np.max(np.abs(vr[m] - vo[m])) <= tol * np.abs(vo[m]).max()
    # size-independent properties: the phasor has unit modulus; rotating back with the transposed matrices and the
    # negated phase vector restores the inputs
    if dtype == np.complex128:
        np.testing.assert_allclose(np.abs(vr[m]), np.abs(vis[m]), rtol=1e-14)
        if not ctr:
            Rt = np.transpose(R, (0, 2, 1)).copy()
            ub = dr.apply_rotation_matrix(ur, field, Rt, rid)
            assert rel_err(ub, uvw) < 1e-15
            vb = dr.apply_phasor(vr, ur, field, freq, -P, rid, ctr, False)
            assert np.max(np.abs(vb[m] - vis[m])) <= 2 * tol * np.abs(vis[m]).max()


def test_field_not_constant_raises():
    from cngi_prototype_b200 import direction_rotate as dr
    uvw = np.zeros((2, 3, 3))
    field = np.array([[1, 1, 2], [1, 1, 1]], dtype=np.int64)
    with pytest.raises(AssertionError, match="constant over baseline"):
        dr.apply_rotation_matrix(uvw, field, np.eye(3)[None].repeat(2, 0), np.array([1, 2]))


def test_dataset_form_and_torch():
    import torch
    from cngi_prototype_b200 import direction_rotate as dr
    d = _golden("direction_rotate_ctr_sp")
    dev = torch.device("cuda")
    vis_ds = {"UVW": torch.as_tensor(d["uvw"], device=dev), "DATA": torch.as_tensor(d["vis"], device=dev),
              "FIELD_ID": torch.as_tensor(d["field"], device=dev), "chan": d["freq_chan"]}
    out = dr.direction_rotate(vis_ds, {"field_id": d["table_ids"], "PHASE_DIR": d["table_dirs"]},
                              {"new_phase_center": d["new_phase_center"]})
    assert out["UVW_ROT"].is_cuda and out["DATA_ROT"].is_cuda and "UVW_ROT" not in vis_ds
    assert rel_err(out["UVW_ROT"].cpu().numpy(), d["uvw_rot"]) < 1e-15
    vr = out["DATA_ROT"].cpu().numpy()
    m = ~np.isnan(vr)
    assert np.max(np.abs(vr[m] - d["vis_rot"][m])) < 1e-6


def test_empty_inputs():
    """Zero integrations / zero channels: nothing is launched, empty outputs come back (ragged channel counts are in
    test_against_oracle: 19 channels = one partial 32-lane chunk)."""
    from cngi_prototype_b200 import direction_rotate as dr
    R, P, rid = np.eye(3)[None], np.zeros((1, 3)), np.array([0])
    v, u = dr.rotate_chunk(np.zeros((0, 4, 3, 2), dtype=np.complex128), np.zeros((0, 4, 3)), np.zeros((0, 4), dtype=np.int64),
                           np.array([1e9, 2e9, 3e9]), R, P, rid)
    assert v.shape == (0, 4, 3, 2) and u.shape == (0, 4, 3)
    v, u = dr.rotate_chunk(np.zeros((2, 4, 0, 2), dtype=np.complex64), np.ones((2, 4, 3)), np.zeros((2, 4), dtype=np.int64),
                           np.zeros(0), R, P, rid)
    assert v.shape == (2, 4, 0, 2) and np.array_equal(u, np.ones((2, 4, 3)))
