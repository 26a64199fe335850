"""CUDA make_pb (cngi_b200_make_pb) against fixtures made by the reference's _airy_disk / _casa_airy_disk
(_make_pb_symmetric.py:26-132) and against the oracle at a larger size.  1e-12 of the peak (the peak is 1)."""
import os

import numpy as np
import pytest

from _util import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["casa_airy", "airy"])
def test_golden(tag):
    from cngi_prototype_b200 import make_pb as mp
    d = np.load(os.path.join(GOLDEN, "pb_%s.npz" % tag))
    gp = dict(image_size=d["image_size"], image_center=d["image_center"], cell_size=d["cell_size"])
    func = mp._casa_airy_disk if tag == "casa_airy" else mp._airy_disk
    for ipower, key in ((2, "pb"), (1, "voltage")):
        got = func(d["freq_chan"], d["pol"], dict(list_dish_diameters=d["dish"], list_blockage_diameters=d["blockage"],
                                                 ipower=ipower), gp).cpu().numpy()
        assert got.shape == d[key].shape
        assert rel_err(got, d[key]) < 1e-12, (key, rel_err(got, d[key]))
        assert np.all(got[12, 10] == 1.0)


def test_make_pb_api_vs_oracle(oracle):
    from cngi_prototype_b200 import make_pb as mp
    freq = np.linspace(345e9, 347e9, 5)
    img = {"chan": freq, "pol": np.array([0, 1, 2])}
    pb_parms = {"list_dish_diameters": [10.7, 6.25, 12.0], "list_blockage_diameters": [0.75, 0.75, 0.0]}
    grid_parms = {"image_size": [300, 257], "cell_size": [0.2, 0.2], "fft_padding": 1.2}
    out = mp.make_pb(img, pb_parms, grid_parms)
    assert "PB" not in img and out["PB"].shape == (300, 257, 5, 3, 3) and list(out["dish_type"]) == [0, 1, 2]
    cell = np.array([-0.2, 0.2]) * np.pi / (3600 * 180)
    ref = oracle.airy_disk(freq, 3, dict(list_dish_diameters=pb_parms["list_dish_diameters"],
                                         list_blockage_diameters=pb_parms["list_blockage_diameters"], ipower=2),
                           dict(image_size=np.array([300, 257]), image_center=np.array([150, 128]), cell_size=cell), casa=True)
    assert rel_err(out["PB"].cpu().numpy(), ref) < 1e-12
    with pytest.raises(AssertionError):
        mp.make_pb(img, {"list_dish_diameters": [10.7], "list_blockage_diameters": [0.75, 0.1]}, grid_parms)
