"""Device-timing probe of cngi_b200_apply_flags on the C2 sample shape and of the zarr -> pinned -> device -> gridder
pipeline (development tool).  Prints one JSON object.

apply_flags: C ABI on preallocated buffers, CUDA events on the launching stream; algorithmic bytes: in place
n * 1 B of flags + elem bytes per flagged element, out of place n * (2 * elem + 1) B.
read_vis: a C1-like store (blosc/zstd chunks, the reference's default compressor) written to a scratch directory,
make_image streamed from it, wall clock (host decode is the bound) next to the in-memory gridding time."""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import _lib, imaging, read_vis as rv, synth  # noqa: E402
from cngi_prototype_b200._devutil import ptr, stream  # noqa: E402
from tools.probe_std_grid import timeit  # noqa: E402

n = 500 * 903 * 128 * 2
L = _lib.lib()
out = {"apply_flags": {}}
flag = (torch.rand(n, device="cuda") < 0.02).to(torch.uint8)
n_flagged = int(flag.sum().item())
for name, dt, kind, eb in (("c64", torch.complex64, _lib.ELEM_C64, 8), ("c128", torch.complex128, _lib.ELEM_C128, 16),
                           ("f32", torch.float32, _lib.ELEM_F32, 4)):
    x = torch.randn(n, dtype=dt, device="cuda")
    y = torch.empty_like(x)
    ms_in, _ = timeit(lambda: _lib.check(L.cngi_b200_apply_flags(ptr(x), ptr(x), ptr(flag), n, kind, None, stream()), "af"), n=9, warm=3)
    ms_out, _ = timeit(lambda: _lib.check(L.cngi_b200_apply_flags(ptr(x), ptr(y), ptr(flag), n, kind, None, stream()), "af"), n=9, warm=3)
    out["apply_flags"][name] = {"samples": n, "flagged": n_flagged,
                                "inplace_ms": ms_in, "inplace_GB/s": (n + n_flagged * eb) / ms_in / 1e6,
                                "copy_ms": ms_out, "copy_GB/s": n * (2 * eb + 1) / ms_out / 1e6}
    del x, y

if "--no-zarr" in sys.argv:          # kernel launches only (the ncu launch list)
    print(json.dumps(out))
    sys.exit(0)
d = synth.config_c1(n_time=600, n_chan=64)
fl = np.random.default_rng(2).random(d["vis"].shape) < 0.02
cell = d["cell"] / imaging.ARCSEC_TO_RAD
gp = {"image_size": [1024, 1024], "cell_size": [cell, cell], "fft_padding": 1.2, "chan_mode": "continuum"}
NAMES = ["DATA", "UVW", "WEIGHT", "FLAG"]
raw_bytes = int(d["vis"].nbytes + d["weight"].nbytes + fl.nbytes + d["uvw"].nbytes)
mem = {"DATA": torch.as_tensor(d["vis"]).cuda(), "UVW": torch.as_tensor(d["uvw"]).cuda(),
       "WEIGHT": torch.as_tensor(d["weight"]).cuda(), "FLAG": torch.as_tensor(fl).cuda().view(torch.uint8),
       "chan": d["freq_chan"]}
imaging.make_image(mem, gp, weight_key="WEIGHT")
torch.cuda.synchronize()
t0 = time.perf_counter()
img2 = imaging.make_image(mem, gp, weight_key="WEIGHT")
torch.cuda.synchronize()
t_mem = time.perf_counter() - t0
out["read_vis"] = {"samples": int(d["vis"].size), "raw_bytes": raw_bytes, "host_cores": os.cpu_count(),
                   "make_image_in_memory_s": t_mem, "time_block": 20, "stores": {}}
# two chunkings of the same samples: 4 and 16 chunk files per variable and 20-integration block
for label, chunks in (("chunks_20x351x16x2", {"time": 20, "chan": 16}), ("chunks_5x351x16x2", {"time": 5, "chan": 16})):
    tmp = tempfile.mkdtemp(prefix="cngi_zarr_")
    try:
        t0 = time.perf_counter()
        store = rv.write_vis(os.path.join(tmp, "c1.vis.zarr"), {"DATA": d["vis"], "UVW": d["uvw"], "WEIGHT": d["weight"],
                                                                "FLAG": fl, "chan": d["freq_chan"]}, chunks=chunks)
        t_write = time.perf_counter() - t0
        nbytes = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(store) for f in fs)
        xds = rv.read_vis(store, partition="xds0").xds0
        res = {"store_bytes": nbytes, "write_s": t_write}
        imaging.make_image(xds, gp, weight_key="WEIGHT", time_chunk=20)      # warm (page cache, cuFFT plan)
        for native in (True, False):
            for workers in (1, 8, 16):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _, blk in xds.iter_device_chunks(NAMES, time_chunk=20, workers=workers, native=native):
                    pass
                torch.cuda.synchronize()
                res["read_only_s_%s_workers%d" % ("native" if native else "python", workers)] = time.perf_counter() - t0
        t0 = time.perf_counter()
        img = imaging.make_image(xds, gp, weight_key="WEIGHT", time_chunk=20)
        torch.cuda.synchronize()
        t_stream = time.perf_counter() - t0
        a, b = img["IMAGE"], img2["IMAGE"].cpu().numpy()
        res.update(make_image_streamed_s=t_stream, streamed_Mvis_per_s=d["vis"].size / t_stream / 1e6,
                   raw_GB_per_s=raw_bytes / t_stream / 1e9,
                   rel_diff_streamed_vs_memory=float(np.abs(a - b).max() / np.abs(b).max()))
        out["read_vis"]["stores"][label] = res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
print(json.dumps(out))
