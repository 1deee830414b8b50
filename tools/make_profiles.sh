#!/bin/bash
# Captures the round's ncu evidence on a B200 (run through gpurun from the repo root):
#   1. launch list with per-launch device time of two DFSPH steps at 2M particles,
#   2. `--set full` captures of the list sweeps (density+first error, correct, error) and of the list builder,
#   3. the same sweeps with SPHK_TILE=1 (TMA-staged tile lists): the kernel-level comparison of the two designs,
# and exports the raw metric pages as CSV so they can be read without a GPU.  Output: gpurun_out/profiles_$TAG/
TAG=${1:-r02}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file $OUT/launches_dfsph_2m.csv \
    python tools/step_probe.py 2m dfsph 2 > $OUT/probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_list -s 20 -c 4 -o $OUT/sweeps_dfsph_2m \
    python tools/step_probe.py 2m dfsph 1 > $OUT/ncu_sweeps.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_list -s 1 -c 1 -o $OUT/build_list_2m \
    python tools/step_probe.py 2m dfsph 1 > $OUT/ncu_build.log 2>&1
SPHK_TILE=1 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tile -s 20 -c 4 -o $OUT/sweeps_tile_dfsph_2m \
    python tools/step_probe.py 2m dfsph 1 > $OUT/ncu_tile.log 2>&1
ncu --set full --clock-control none -k regex:"k_gather|k_hash_snapshot|k_cell_start|DeviceRadixSortOnesweep" -s 21 -c 7 -o $OUT/search_2m \
    python tools/step_probe.py 2m dfsph 1 > $OUT/ncu_search.log 2>&1
for f in sweeps_dfsph_2m build_list_2m sweeps_tile_dfsph_2m search_2m; do
  ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/$f.raw.csv > $OUT/$f.summary.txt 2>/dev/null
done
python tools/agg_launches.py $OUT/launches_dfsph_2m.csv > $OUT/launches_dfsph_2m.summary.txt 2>/dev/null
rm -f $OUT/*.ncu-rep.tmp
ls -la $OUT
