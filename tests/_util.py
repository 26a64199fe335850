"""Shared helpers for the parity tests (golden fixture loading, comparison metrics)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    gp = dict(chan_mode=str(d["gp_chan_mode"]), image_size_padded=d["gp_image_size_padded"].astype(np.int64),
              cell_size=d["gp_cell_size"], support=int(d["gp_support"]), do_psf=bool(d["gp_do_psf"]),
              complex_grid=bool(d["gp_complex_grid"]), do_imaging_weight=bool(d["gp_do_imaging_weight"]))
    os_ = d["gp_oversampling"]
    gp["oversampling"] = int(os_) if os_.ndim == 0 else os_.astype(np.int64)
    if "gp_field_id" in d:
        gp["field_id"] = d["gp_field_id"].astype(np.int64)
    return d, gp


def rel_err(a, b):
    """max|a-b| / max|b| -- the metric the reference's notebooks use."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.abs(b))
    if scale == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / scale)


def same_support(a, b):
    """Bit-exact check of which cells were touched (cell indexing + masking)."""
    return np.array_equal(np.asarray(a) != 0, np.asarray(b) != 0)
