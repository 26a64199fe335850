#!/bin/bash
# Run ON THE GPU BOX (gpurun -- tools/collect_profiles_r02.sh): regenerates the round-2 artefacts under gpurun_out/;
# tools/make_profiles_r02.py (run in the container afterwards) turns them into profiles/r02_*.
set -x
O=gpurun_out
mkdir -p $O
B="--no-cpu-baseline --no-e2e --no-cube --no-extras --no-parity --no-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 $B > $O/r02_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:std_grid_window -c 1 -f -o $O/r02_window \
    python bench.py --steps 1 --warmup 3 $B > $O/r02_ncu_window.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:std_grid_window -c 1 -f -o $O/r02_window_iw \
    python bench.py --steps 1 --warmup 3 $B --fuse-weights > $O/r02_ncu_window_iw.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:bluestein --launch-skip 2 -c 2 -f -o $O/r02_bluestein \
    python tools/probe_fft_one.py 9830 8192 > $O/r02_ncu_blu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:std_grid_window --launch-skip 1 -c 1 -f -o $O/r02_window_cube \
    python tools/probe_cube_chunk.py 2000 > $O/r02_ncu_cube.log 2>&1
if [ -z "$SKIP_APERTURE" ]; then   # (kernel unchanged since the first collection of the round: SKIP_APERTURE=1 keeps those files)
ncu --set full --import-source on --clock-control none -k regex:aperture_track -c 1 -f -o $O/r02_aperture \
    python tools/probe_aperture.py > $O/r02_ncu_aperture.log 2>&1
CNGI_APERTURE_BULK=1 ncu --set full --import-source on --clock-control none -k regex:aperture_track -c 1 -f -o $O/r02_aperture_bulk \
    python tools/probe_aperture.py > $O/r02_ncu_aperture_bulk.log 2>&1
fi
# the imaging-weight kernels, both generations (tools/probe_iw_one.py runs the general kernels, then the fast ones)
ncu --set full --import-source on --clock-control none -k regex:"iw_grid|iw_degrid" -c 4 -f -o $O/r02_iw \
    python tools/probe_iw_one.py > $O/r02_ncu_iw.log 2>&1
python tools/probe_fft.py 2> $O/r02_fft.err | tail -1 > $O/r02_fft.json
python tools/probe_fused_weights.py f32 f64 2> $O/r02_fused_weights.err | tail -1 > $O/r02_fused_weights.json
python tools/bench_rows.py > $O/r02_rows.json 2> $O/r02_rows.err
python tools/make_profiles_r02.py > $O/r02_make_profiles.log 2>&1
rm -f $O/*.ncu-rep          # summarised above; together they exceed the 64 MiB gpurun copies back
tail -c 300 $O/r02_rows.json
