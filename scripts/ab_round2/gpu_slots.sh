#!/bin/bash
# two vs three overlapped batch pipelines; ncu of the half-warp traceback at 128 x 128 tiles
O=gpurun_out/${1:-slots}
mkdir -p $O
for n in 2 3; do
  SEQALK_SLOTS=$n python scripts/ab_lib.py cfg4 2 2>&1 | sed "s/^/slots=$n /" | tee -a $O/ab.txt
  SEQALK_SLOTS=$n python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/slots=$n /" | tee -a $O/span.txt
done
SEQALK_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:traceback_tiles -c 1 \
    -o $O/tb_half128 python scripts/ab_lib.py cfg3 1 > $O/ncu_tb.log 2>&1
ncu -i $O/tb_half128.ncu-rep --page raw --csv > $O/tb_half128_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/tb_half128_raw.csv > $O/tb_half128_summary.txt 2>&1
ncu -i $O/tb_half128.ncu-rep --page source --csv > $O/tb_half128_source.csv 2>/dev/null
python scripts/sass_blocks.py $O/tb_half128_source.csv >> $O/tb_half128_summary.txt 2>&1
cat $O/tb_half128_summary.txt
