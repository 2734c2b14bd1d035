#!/bin/bash
# The round's profiling recipe (run on the GPU box through gpurun): plain runs first (each must exit 0), then ncu.
#   bash tools/profile_round.sh <tag>
set -u
T=${1:-rX}
O=gpurun_out
python tools/prof_step.py --batch 2048 --steps 2 > $O/${T}_prof_plain.log 2>&1 || { echo "prof_step failed"; exit 1; }
python bench.py --steps 2 --warmup 1 > $O/${T}_launch_plain.log 2>&1 || { echo "bench failed"; exit 1; }
python tools/bench_render.py --batch 64 --steps 1 --warmup 1 > $O/${T}_render_plain.log 2>&1 || { echo "render failed"; exit 1; }
# (1) launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 > $O/${T}_launch_ncu.log 2>&1
# (2) full capture of the library's kernels, one step at 2048 samples
ncu --set full --clock-control none --import-source on -k "regex:pose_|blend_|lbs_|mask_|seg_|split3" -o $O/${T}_prof_step2048 -f \
    python tools/prof_step.py --batch 2048 --steps 1 > $O/${T}_prof_ncu.log 2>&1
# (3) full capture of the visualiser's two kernels
ncu --set full --clock-control none --import-source on -k regex:render_ -c 2 -o $O/${T}_prof_render -f \
    python tools/bench_render.py --batch 64 --steps 1 --warmup 0 > $O/${T}_render_ncu.log 2>&1
ls -la $O/${T}_*
