#!/bin/bash
# Captures the round's ncu evidence on a B200 (run through gpurun from the repo root):
#   1. launch list with per-launch device time of two DFSPH steps at 2M particles,
#   2. `--set full` captures of the list sweeps (first three list launches of a step) and of the list builder,
# and exports the raw metric pages as CSV so they can be read without a GPU.  Output: gpurun_out/profiles_$TAG/
TAG=${1:-r01}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file $OUT/launches_dfsph_2m.csv \
    python tools/step_probe.py 2m dfsph 2 > $OUT/probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_list -s 21 -c 4 -o $OUT/sweeps_dfsph_2m \
    python tools/step_probe.py 2m dfsph 1 > $OUT/ncu_sweeps.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_list -s 1 -c 1 -o $OUT/build_list_2m \
    python tools/step_probe.py 2m dfsph 1 > $OUT/ncu_build.log 2>&1
for f in sweeps_dfsph_2m build_list_2m; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null
done
ls -la $OUT
