#!/bin/bash
# round-2 checkpoint on one B200: GPU tests, default bench line, ncu of the packed fill (RP = 32)
TAG=${1:-r2p}
O=gpurun_out/$TAG
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
python bench.py --steps 1 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python scripts/ab_lib.py cfg3 2 > $O/ab_cfg3.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fill16_kernel -c 1 \
    -o $O/fill16_rp32 python scripts/ab_lib.py cfg3 1 > $O/ncu_fill16.log 2>&1
echo "ncu rc=$?"
ncu -i $O/fill16_rp32.ncu-rep --page raw --csv > $O/fill16_rp32_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/fill16_rp32_raw.csv > $O/fill16_rp32_summary.txt 2>&1
cat $O/ab_cfg3.txt; head -c 1500 $O/bench_n1.json
