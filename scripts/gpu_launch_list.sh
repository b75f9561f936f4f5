#!/bin/bash
# Runs on the GPU box: the default bench command plain, then the same command under ncu's
# duration-only pass (one ncu use per gpurun call).  usage: bash scripts/gpu_launch_list.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_list_$TAG.json 2> gpurun_out/plain_list_$TAG.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
