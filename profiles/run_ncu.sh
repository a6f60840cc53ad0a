#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for this repo (run under gpurun).
# usage: profiles/run_ncu.sh <tag> [extra bench args]  -> gpurun_out/{plain,launches,prof_build,prof_edit}_<tag>.*
# Every ncu pass runs only after the same command exited 0 without ncu; numbers printed under ncu are not bench values.
set -u
TAG=${1:-r2}; shift || true
ARGS="--config 2 --steps 2 --warmup 3 --genome-len 1000000 --no-cpu-baseline --no-roof $*"
mkdir -p gpurun_out
python bench.py $ARGS --separate > gpurun_out/plain_separate_$TAG.json 2> gpurun_out/plain_separate_$TAG.log
python bench.py $ARGS > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu1_$TAG.log 2>&1
python bench.py $ARGS > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:build_filters_levels -s 4 -c 1 \
    -o gpurun_out/prof_build_$TAG -f python bench.py $ARGS > gpurun_out/ncu2_$TAG.log 2>&1
python bench.py $ARGS > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:edit_kernel -s 3 -c 1 \
    -o gpurun_out/prof_edit_$TAG -f python bench.py $ARGS > gpurun_out/ncu3_$TAG.log 2>&1
tail -3 gpurun_out/ncu2_$TAG.log gpurun_out/ncu3_$TAG.log
cat gpurun_out/plain_$TAG.json
