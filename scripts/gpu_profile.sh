#!/bin/bash
# Runs on the GPU box: launch list of the default bench command + full ncu captures of the two
# dominant kernels (each after the same command ran plain and exited 0).
# usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
LIST_CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
FULL_CMD="python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline"
$LIST_CMD > gpurun_out/plain_list_$TAG.json 2> gpurun_out/plain_list_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $LIST_CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
$FULL_CMD > gpurun_out/plain_full_$TAG.json 2> gpurun_out/plain_full_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s 2 -c 1 \
    -o gpurun_out/fill_$TAG $FULL_CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:traceback_tiles -s 2 -c 1 \
    -o gpurun_out/tbtiles_$TAG $FULL_CMD > gpurun_out/ncu_full_tb_$TAG.log 2>&1
echo "tb capture rc=$?"
