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
python tools/red_peak.py > $O/r01_red_peak.json 2> $O/r01_red_peak.err
python bench.py > $O/r01_bench_line.json 2> $O/r01_bench.err
python tools/bench_rows.py > $O/r01_rows.json 2> $O/r01_rows.err
tail -c 400 $O/r01_bench_line.json
