#!/bin/bash
# Builds a tuning variant of the library: tools/build_variant.sh <name> <extra nvcc flags...>
# -> gpurun_variants/libcngi_b200_<name>.so ; select it at run time with CNGI_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
objs=""
for f in cngi_prototype_b200/csrc/*.cu; do
  o=variants/$(basename ${f%.cu})_$name.o
  if [ "$(basename $f)" = "standard_grid.cu" ] || [ ! -f variants/$(basename ${f%.cu})_base.o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 -I include "$@" -c $f -o $o -Xptxas -v 2> variants/$(basename ${f%.cu})_$name.log
  else
    o=variants/$(basename ${f%.cu})_base.o
  fi
  objs="$objs $o"
done
nvcc -shared -o variants/libcngi_b200_$name.so $objs -L /usr/local/cuda/lib64 -lcufft -lcudart -Xlinker -rpath,/usr/local/cuda/lib64
grep -A2 "track_kernelIfLb1ELi7ELi2ELi128" variants/standard_grid_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
grep -A2 "track_kernelIfLb1ELi7ELi2ELi256" variants/standard_grid_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
