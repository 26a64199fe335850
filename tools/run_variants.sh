#!/bin/bash
# run on the GPU box: tools/run_variants.sh <algo> <variant names...> -- times each tuning variant on config C2 (fp32)
algo=$1; shift
for v in "$@"; do
  echo "== $v"
  CNGI_B200_LIB=$PWD/variants/libcngi_b200_$v.so timeout 300 python tools/probe_std_grid.py --config c2 --precs f32 --algos $algo 2>&1 | grep ms
done
