#!/bin/bash
O=gpurun_out/${1:-nculong}
mkdir -p $O
SEQALK_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:traceback_tiles_long -c 1 \
    -o $O/tb_long python scripts/ab_lib.py cfg5 1 > $O/ncu.log 2>&1
ncu -i $O/tb_long.ncu-rep --page raw --csv > $O/tb_long_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/tb_long_raw.csv > $O/tb_long_summary.txt 2>&1
ncu -i $O/tb_long.ncu-rep --page source --csv > $O/tb_long_source.csv 2>/dev/null
python scripts/sass_blocks.py $O/tb_long_source.csv 0.02 >> $O/tb_long_summary.txt 2>&1
cat $O/tb_long_summary.txt
