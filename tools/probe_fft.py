"""grid -> image time per plane at the padded sizes of power-of-two images: shared-memory Bluestein passes vs cuFFT."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import _fft  # noqa: E402
from probe_std_grid import timeit  # noqa: E402

out = {}
for n_pad, n_img, planes in ((9830, 8192, 4), (4915, 4096, 4), (1228, 1024, 16), (10240, 8192, 4)):
    g = torch.view_as_complex(torch.randn((planes // 2, 2, n_pad, n_pad, 2), dtype=torch.float32, device="cuda"))
    row = {}
    for knob in ("1", "0"):
        os.environ["CNGI_FFT_BLUESTEIN"] = knob
        _fft._plans.clear()
        img = _fft.grid_to_image(g, (n_img, n_img))
        row["bluestein" if knob == "1" else "cufft"] = round(timeit(lambda: _fft.grid_to_image(g, (n_img, n_img)))[0] / planes, 3)
        if knob == "1":
            keep = img.clone()
        else:
            row["max_rel_diff"] = float((img - keep).abs().max() / img.abs().max())
    out["%d^2 -> %d^2" % (n_pad, n_img)] = row
    del g, img, keep
    _fft._plans.clear()
    torch.cuda.empty_cache()
print(json.dumps({"ms_per_plane_complex64_incl_crop_pass": out}))
