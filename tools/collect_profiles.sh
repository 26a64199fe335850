#!/bin/bash
# Run ON THE GPU BOX (gpurun -- tools/collect_profiles.sh): regenerates every artefact profiles/ is built from into
# gpurun_out/.  tools/make_profiles.py (run in the container afterwards) turns them into profiles/r01_*.
set -x
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r01_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:std_grid_window -c 1 -f -o $O/r01_window \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r01_ncu_window.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:iw_ -c 8 -f -o $O/r01_iw \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r01_ncu_iw.log 2>&1
# section 8(f) rows: the fused image+psf launch is the 15th window-kernel launch of probe_fused.py (7 image + 7 psf first)
ncu --set full --import-source on --clock-control none -k regex:std_grid_window --launch-skip 14 -c 1 -f -o $O/r01_fused \
    python tools/probe_fused.py > $O/r01_ncu_fused.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:dr_phasor -c 1 -f -o $O/r01_dr \
    python tools/probe_direction_rotate.py > $O/r01_ncu_dr.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"dr_|gcf_" -c 60 --csv \
    --log-file $O/r01_next_launches.csv python tools/probe_gcf.py > $O/r01_ncu_gcf.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:aperture_track -c 1 -f -o $O/r01_aperture \
    python tools/probe_aperture.py > $O/r01_ncu_aperture.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:std_degrid_window -c 1 -f -o $O/r01_degrid \
    python tools/probe_degrid.py > $O/r01_ncu_degrid.log 2>&1
python tools/probe_fused.py 2> $O/r01_fused.err | tail -1 > $O/r01_fused.json
python tools/probe_direction_rotate.py 2> $O/r01_dr.err | tail -1 > $O/r01_direction_rotate.json
python tools/probe_gcf.py --cpu 2> $O/r01_gcf.err | tail -1 > $O/r01_gcf.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:apply_flags -c 80 --csv \
    --log-file $O/r01_apply_flags_launches.csv python tools/probe_apply_flags.py --no-zarr > $O/r01_ncu_apply_flags.log 2>&1
python tools/probe_apply_flags.py > $O/r01_apply_flags.json 2> $O/r01_apply_flags.err
python tools/red_peak.py > $O/r01_red_peak.json 2> $O/r01_red_peak.err
python bench.py > $O/r01_bench_line.json 2> $O/r01_bench.err
python tools/bench_rows.py > $O/r01_rows.json 2> $O/r01_rows.err
tail -c 400 $O/r01_bench_line.json
