#!/bin/bash
# Builds a tuning variant of the library: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>
# Recompiles only csrc/<file.cu> with the extra flags and links it with the objects of the regular build
# (python -m cngi_prototype_b200.build first) -> variants/libcngi_b200_<name>.so ; select it at run time with
# CNGI_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
name=$1; file=$2; shift; shift
mkdir -p variants
objs=""
for f in cngi_prototype_b200/csrc/*.cu; do
  if [ "$(basename $f)" = "$file" ]; then
    o=variants/$(basename ${f%.cu})_$name.o
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 -I include "$@" -c $f -o $o -Xptxas -v 2> variants/$(basename ${f%.cu})_$name.log
  else
    o=${f%.cu}.o
  fi
  objs="$objs $o"
done
nvcc -shared -o variants/libcngi_b200_$name.so $objs -L /usr/local/cuda/lib64 -lcufft -lcudart -Xlinker -rpath,/usr/local/cuda/lib64
grep -A2 "kernelIfLb1ELi7ELi2ELi128" variants/${file%.cu}_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
