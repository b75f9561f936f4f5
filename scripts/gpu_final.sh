#!/bin/bash
# Final single-GPU record of a round: GPU tests, the default bench line, the ncu launch list of the
# bench command, full ncu captures of the two dominant kernels (each after its command ran plain).
# usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r02_final}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
python bench.py --steps 2 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
LIST_CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary"
$LIST_CMD > $O/plain_list.json 2> $O/plain_list.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $O/launches.csv $LIST_CMD > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
JOB_CMD="python scripts/ab_lib.py cfg4 1"
SEQALK_OVERLAP=0 $JOB_CMD > $O/plain_job.txt 2>&1 || exit 1
for k in fill16_all traceback_tiles_half; do
  SEQALK_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 \
      -o $O/$k $JOB_CMD > $O/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  ncu -i $O/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $O/${k}_raw.csv > $O/${k}_summary.txt 2>&1
  ncu -i $O/$k.ncu-rep --page source --csv > $O/${k}_source.csv 2>/dev/null
  python scripts/sass_blocks.py $O/${k}_source.csv 0.02 >> $O/${k}_summary.txt 2>&1
done
cat $O/plain_job.txt; head -c 400 $O/bench_n1.json
