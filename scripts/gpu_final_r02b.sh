#!/bin/bash
# Final single-GPU record of round 2 (second half): smoke, GPU tests, the default bench line, the ncu launch
# list of the bench command, one full ncu capture of the dominant kernel (each ncu pass after its command ran plain).
O=gpurun_out/${1:-r02b_final}
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
LIST_CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary"
$LIST_CMD > $O/plain_list.json 2> $O/plain_list.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $O/launches.csv $LIST_CMD > $O/ncu_list.log 2>&1
echo "launch list rc=$?"
JOB_CMD="python scripts/ab_lib.py cfg4 1"
SEQALK_OVERLAP=0 $JOB_CMD > $O/plain_job.txt 2>&1 || exit 1
for k in fill16_all; do
  SEQALK_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 \
      -o $O/$k $JOB_CMD > $O/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
  ncu -i $O/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $O/${k}_raw.csv > $O/${k}_summary.txt 2>&1
  ncu -i $O/$k.ncu-rep --page source --csv > $O/${k}_source.csv 2>/dev/null
  python scripts/sass_blocks.py $O/${k}_source.csv 0.02 >> $O/${k}_summary.txt 2>&1
  rm -f $O/${k}_source.csv $O/${k}_raw.csv
done
cat $O/plain_job.txt; head -c 600 $O/bench_n1.json
