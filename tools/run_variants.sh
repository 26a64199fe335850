#!/bin/bash
# run on the GPU box: times each tuning variant of the track kernel on config C2 (fp32 continuum + cube)
for v in "$@"; do
  lib=variants/libcngi_b200_${v%%:*}.so; blk=${v##*:}
  echo "== $v"
  CNGI_B200_LIB=$PWD/$lib CNGI_TRACK_BLOCK=$blk timeout 300 python tools/probe_std_grid.py --config c2 --precs f32 --algos 2 2>&1 | grep ms
done
