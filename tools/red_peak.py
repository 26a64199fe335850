"""Measures the REDG.F32x2 reduction ceiling of the box in 32-byte sectors per second (development tool; bench.py
runs the same probe for its `atomic_roofline`).

    python tools/red_peak.py > profiles/r01_red_peak.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import _lib  # noqa: E402
from cngi_prototype_b200._devutil import ptr, stream  # noqa: E402

SECTORS_PER_INSTR = {0: 32, 1: 8, 2: 8}
NAMES = {0: "scattered (one sector per lane)", 1: "8 lanes x 8 B contiguous (64 B)", 2: "32 lanes x 8 B contiguous (256 B)"}


def red_rate(n_cells, pattern, blocks=148 * 16, per_thread=256, reps=3):
    """Gsectors/s of pattern `pattern` into a buffer of n_cells 8-byte cells."""
    L = _lib.lib()
    buf = torch.zeros(n_cells, dtype=torch.complex64, device="cuda")
    best = float("inf")
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.cngi_b200_microbench_red(ptr(buf), n_cells, pattern, blocks, per_thread, stream()), "microbench_red")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    warp_instr = blocks * 8 * per_thread
    return warp_instr * SECTORS_PER_INSTR[pattern] / (best * 1e-3) / 1e9, best


def main():
    _lib.require_device()
    out = []
    for mb in (32, 268, 4096, 34000):   # L2 resident, the C2 continuum grid, > L2, the C2 cube
        n_cells = mb * 1000 * 1000 // 8
        for pattern in (0, 1, 2):
            g, ms = red_rate(n_cells, pattern)
            out.append({"footprint_mb": mb, "pattern": NAMES[pattern], "gsectors_per_s": round(g, 2), "ms": round(ms, 3)})
            print(json.dumps(out[-1]), flush=True)
    # the alternative design in isolation: shared-memory fp atomics (ATOMS.CAST.SPIN on sm_100a)
    L = _lib.lib()
    sink = torch.zeros(1024, dtype=torch.complex64, device="cuda")
    blocks, per_thread, best = 148 * 8, 64, float("inf")
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.cngi_b200_microbench_smem_atomics(ptr(sink), blocks, per_thread, stream()), "microbench_smem_atomics")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    smem = {"design": "32x32 shared-memory subgrid per block, atomicAdd(float) x2 per complex tap, no loads / index math",
            "complex_tap_updates_per_s": blocks * 256 * per_thread * 49 / (best * 1e-3), "ms": round(best, 3)}
    print(json.dumps(smem), flush=True)
    print(json.dumps({"red_peak": out, "smem_atomics": smem}))


if __name__ == "__main__":
    main()
