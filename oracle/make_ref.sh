#!/bin/sh
# oracle/make_ref.sh -- TEST / BENCH INFRASTRUCTURE, never part of the product.
#
# Stages the reference's OWN numba gridding loops so that they can be timed on the GPU box's host cores
# (bench.py `cpu_baseline.kind = "reference"` and `--impl reference`), where /root/reference does not exist.
# The reference is pure Python + numba, so "compiling it from its own sources where they lie" means: place the
# unmodified files, in their own package-relative layout, under the git-ignored oracle/_ref/ (it travels to the
# GPU box with the gpurun snapshot like a built .so, and never enters history), and let numba JIT them there.
# Only the files on the hot path are staged:
#     ngcasa/imaging/_imaging_utils/{_standard_grid,_aperture_grid,_gridding_convolutional_kernels}.py
#     cngi/_utils/_constants.py                (imported by _aperture_grid.py:19)
# The package __init__ files are written EMPTY here (the reference's own import xarray/dask, absent in this image).
# oracle/ref_loader.py loads them by path with CNGI_REFERENCE_ROOT=oracle/_ref.
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/ngcasa/imaging/_imaging_utils/_standard_grid.py" ]; then
    echo "make_ref.sh: reference tree not found at $REF (keeping whatever is in $OUT)" >&2
    exit 0
fi
rm -rf "$OUT"
mkdir -p "$OUT/ngcasa/imaging/_imaging_utils" "$OUT/cngi/_utils"
for f in _standard_grid.py _aperture_grid.py _gridding_convolutional_kernels.py; do
    cp "$REF/ngcasa/imaging/_imaging_utils/$f" "$OUT/ngcasa/imaging/_imaging_utils/$f"
done
cp "$REF/cngi/_utils/_constants.py" "$OUT/cngi/_utils/_constants.py"
: > "$OUT/cngi/__init__.py"
: > "$OUT/cngi/_utils/__init__.py"
( cd "$OUT" && find . -name '*.py' -size +0 | sort | xargs sha256sum ) > "$OUT/MANIFEST.sha256"
echo "staged $(wc -l < "$OUT/MANIFEST.sha256") reference files under $OUT"
