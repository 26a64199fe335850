"""direction_rotate: rotate uvw to a new phase centre and phase-rotate the visibilities (joint mosaics).

Mirrors /root/reference/ngcasa/imaging/direction_rotate.py on array inputs (the reference's entry point takes an
xarray mxds; xarray is not in this image, so the dataset here is a mapping, like cngi_prototype_b200/imaging.py):

  calc_rotation_mats     :127-175   host numpy (n_field 3x3 matrices; nothing to accelerate)
  apply_rotation_matrix  :190-213   -> cngi_b200_direction_rotate (uvw only)
  apply_phasor           :217-248   -> cngi_b200_direction_rotate
  direction_rotate       :29-124    -> direction_rotate(vis_dataset, field_dataset, rotation_parms, sel_parms)

numpy in -> numpy out, torch-CUDA in -> torch-CUDA out; no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._devutil import torch, is_torch, precision_of, torch_dtypes, device_of, Uploader, ptr, stream, back


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _directional_cosine(phase_center_in_radians):
    """(RA, DEC) -> direction cosines (direction_rotate.py:177-188)."""
    ra, dec = phase_center_in_radians[0], phase_center_in_radians[1]
    return np.array([np.cos(ra) * np.cos(dec), np.sin(ra) * np.cos(dec), np.sin(dec)])


def calc_rotation_mats(field_id_of_samples, field_ids, field_phase_dir, rotation_parms):
    """Per-field uvw rotation matrix and phase-rotation vector (direction_rotate.py:127-175).

    field_id_of_samples: vis_dataset.FIELD_ID (n_time, n_baseline); field_ids / field_phase_dir: the FIELD table
    (ids and (n, 2) RA/DEC in radians).  The reference composes scipy intrinsic Euler rotations 'XZ' and 'ZX';
    written out here as Rx(pi/2 - dec) Rz(pi/2 - ra) and Rz(ra_f - pi/2) Rx(dec_f - pi/2).
    Returns uvw_rotmat (n_field, 3, 3), phase_rotation (n_field, 3), rot_field_id (n_field)."""
    ra, dec = float(rotation_parms["new_phase_center"][0]), float(rotation_parms["new_phase_center"][1])
    rot_new = _rot_x(np.pi / 2 - dec) @ _rot_z(-ra + np.pi / 2)
    cos_new = _directional_cosine((ra, dec))
    f = field_id_of_samples.cpu().numpy() if is_torch(field_id_of_samples) else np.asarray(field_id_of_samples)
    ids = np.unique(f)
    ids = ids[ids > -1]
    table_ids = [int(i) for i in np.asarray(field_ids)]
    dirs = np.asarray(field_phase_dir, dtype=np.float64)
    uvw_rotmat = np.zeros((len(ids), 3, 3))
    phase_rotation = np.zeros((len(ids), 3))
    for i, fid in enumerate(ids):
        pc = dirs[table_ids.index(int(fid))]
        rot_field = _rot_z(-np.pi / 2 + pc[0]) @ _rot_x(pc[1] - np.pi / 2)
        uvw_rotmat[i] = (rot_new @ rot_field).T
        if rotation_parms.get("common_tangent_reprojection", True):
            uvw_rotmat[i, 2, 0:2] = 0.0      # joint mosaics: FTMachine::girarUVW's common tangent (:166-167)
        phase_rotation[i] = rot_new @ (cos_new - _directional_cosine(pc))
    return uvw_rotmat, phase_rotation, ids.astype(np.int64)


def _rotate(vis_data, uvw, field_id, freq_chan, uvw_rotmat, phase_rotation, rot_field_id,
            common_tangent_reprojection, single_precision, want_uvw):
    L = _lib.lib()
    like_torch = is_torch(uvw) or is_torch(vis_data)
    dev = device_of(vis_data, uvw)
    up = Uploader(dev)
    uvw_t = up(uvw, torch.float64)
    n_time, n_baseline = int(uvw_t.shape[0]), int(uvw_t.shape[1])
    fld = field_id
    if not is_torch(fld):
        fld = np.asarray(fld)
    fld = up(fld.reshape(n_time, n_baseline), torch.int64)
    a = _lib.DirectionRotateArgs()
    vis_rot = None
    if vis_data is not None:
        precision = precision_of(vis_data)
        _, cdt = torch_dtypes(precision)
        vis_t = up(vis_data, cdt)
        vis_rot = torch.empty_like(vis_t)
        a.n_chan, a.n_pol = int(vis_t.shape[2]), int(vis_t.shape[3])
        a.vis, a.vis_rot, a.precision = ptr(vis_t), ptr(vis_rot), precision
        a.freq_chan = ptr(up(np.asarray(freq_chan).reshape(-1) if not is_torch(freq_chan) else freq_chan.reshape(-1),
                             torch.float64))
    else:
        a.n_chan, a.n_pol, a.precision = 0, 1, _lib.F64
    uvw_rot = torch.empty_like(uvw_t) if want_uvw else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    a.n_time, a.n_baseline = n_time, n_baseline
    a.uvw, a.uvw_rot, a.field = ptr(uvw_t), ptr(uvw_rot), ptr(fld)
    a.uvw_rotmat = ptr(up(uvw_rotmat, torch.float64))
    a.phase_rotation = ptr(up(phase_rotation, torch.float64))
    ids = up(rot_field_id, torch.int64)
    a.rot_field_id, a.n_field, a.status = ptr(ids), int(ids.numel()), ptr(status)
    a.common_tangent_reprojection = int(bool(common_tangent_reprojection))
    a.single_precision = int(bool(single_precision))
    with torch.cuda.device(dev):
        _lib.check(L.cngi_b200_direction_rotate(C.byref(a), stream()), "cngi_b200_direction_rotate")
    # the reference asserts inside the loop (:200,:227); same message, raised after the launch
    assert int(status.item()) == 0, "direction_rotate only supports xds where field_id remains constant over baseline."
    return (None if vis_rot is None else back(vis_rot, like_torch),
            None if uvw_rot is None else back(uvw_rot, like_torch))


def apply_rotation_matrix(uvw, field_id, uvw_rotmat, rot_field_id):
    """uvw (n_time, n_baseline, 3) -> rotated uvw.  field_id is (n_time, n_baseline[, 1]) as in the reference."""
    dummy_phase = np.zeros((len(uvw_rotmat), 3))
    return _rotate(None, uvw, field_id, None, uvw_rotmat, dummy_phase, rot_field_id, True, False, True)[1]


def apply_phasor(vis_data, uvw, field_id, freq_chan, phase_rotation, rot_field_id, common_tangent_reprojection,
                 single_precision):
    """vis * exp(2 pi i d f / c) with d from the ALREADY ROTATED uvw (the reference passes uvw_rot[..., None]).
    uvw: (n_time, n_baseline, 3[, 1]); field_id (n_time, n_baseline[, 1, 1]); freq_chan any shape with n_chan values."""
    uvw3 = uvw.reshape(uvw.shape[0], uvw.shape[1], 3)
    identity = np.broadcast_to(np.eye(3), (len(phase_rotation), 3, 3)).copy()
    return _rotate(vis_data, uvw3, field_id, freq_chan, identity, phase_rotation, rot_field_id,
                   common_tangent_reprojection, single_precision, False)[0]


def rotate_chunk(vis_data, uvw, field_id, freq_chan, uvw_rotmat, phase_rotation, rot_field_id,
                 common_tangent_reprojection=True, single_precision=True):
    """Both steps in one call (one pass over uvw for the phase, one for the rotated uvw): returns (vis_rot, uvw_rot)."""
    return _rotate(vis_data, uvw, field_id, freq_chan, uvw_rotmat, phase_rotation, rot_field_id,
                   common_tangent_reprojection, single_precision, True)


def direction_rotate(vis_dataset, field_dataset, rotation_parms, sel_parms=None):
    """Mapping-dataset form of the reference entry point (:29-124).  vis_dataset: {'UVW','DATA','FIELD_ID','chan'};
    field_dataset: {'field_id','PHASE_DIR' (n_field, 2)}.  Adds sel_parms['data_group_out'] names (default
    'UVW_ROT' / 'DATA_ROT') to a shallow copy of vis_dataset and returns it; inputs are not modified."""
    sel = dict(sel_parms or {})
    din = {"uvw": "UVW", "data": "DATA", **sel.get("data_group_in", {})}
    dout = {"uvw": "UVW_ROT", "data": "DATA_ROT", **sel.get("data_group_out", {})}
    parms = {"common_tangent_reprojection": True, "single_precision": True, **rotation_parms}
    assert len(parms["new_phase_center"]) == 2, "######### ERROR: rotation_parms checking failed"
    rotmat, phase_rot, ids = calc_rotation_mats(vis_dataset["FIELD_ID"], field_dataset["field_id"],
                                                field_dataset["PHASE_DIR"], parms)
    vis_rot, uvw_rot = rotate_chunk(vis_dataset[din["data"]], vis_dataset[din["uvw"]], vis_dataset["FIELD_ID"],
                                    vis_dataset["chan"], rotmat, phase_rot, ids,
                                    parms["common_tangent_reprojection"], parms["single_precision"])
    out = dict(vis_dataset)
    out[dout["uvw"]], out[dout["data"]] = uvw_rot, vis_rot
    return out
