"""Loads the UNMODIFIED reference numba gridders by file path.

TEST / BENCH INFRASTRUCTURE.  In this container the files are read where they lie under /root/reference
(tests/golden/make_golden.py pins the C restatement with them).  /root/reference does not exist on the GPU box:
there the copy staged by oracle/make_ref.sh under the git-ignored oracle/_ref/ is loaded instead -- only by
bench.py's CPU legs (cpu_baseline kind "reference", `--impl reference`) and by tests that skip when it is absent.
Recipe from SURVEY.md Appendix A: register modules in sys.modules before exec (numba
cache=True re-imports by name), writable NUMBA_CACHE_DIR, and the removed ``np.int`` alias.
"""
import importlib.util
import os
import sys

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    env = os.environ.get("CNGI_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(root, "ngcasa", "imaging", "_imaging_utils", "_standard_grid.py")):
            return root
    return "/root/reference"


REF_ROOT = _pick_root()
_UTILS = os.path.join(REF_ROOT, "ngcasa", "imaging", "_imaging_utils")


def available():
    return os.path.isfile(os.path.join(_UTILS, "_standard_grid.py"))


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns (standard_grid_module, aperture_grid_module, ps_kernels_module)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_cngi_ref")
    if not hasattr(np, "int"):
        np.int = int  # only alias the wrappers need (_standard_grid.py:153)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)  # for cngi._utils._constants (_aperture_grid.py:19)
    sg = _load("ref_standard_grid", os.path.join(_UTILS, "_standard_grid.py"))
    ag = _load("ref_aperture_grid", os.path.join(_UTILS, "_aperture_grid.py"))
    ck = _load("ref_ps_kernels", os.path.join(_UTILS, "_gridding_convolutional_kernels.py"))
    return sg, ag, ck
