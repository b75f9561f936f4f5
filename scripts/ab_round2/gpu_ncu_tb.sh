#!/bin/bash
# ncu --set full of the tile traceback (cfg3), half-warp kernel and warp kernel
O=gpurun_out/${1:-ncutb}
mkdir -p $O
for h in 1 0; do
  SEQALK_TB_HALF=$h SEQALK_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:traceback_tiles -c 1 \
    -o $O/tb_half$h python scripts/ab_lib.py cfg3 1 > $O/ncu_half$h.log 2>&1
  ncu -i $O/tb_half$h.ncu-rep --page raw --csv > $O/tb_half${h}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $O/tb_half${h}_raw.csv > $O/tb_half${h}_summary.txt 2>&1
  ncu -i $O/tb_half$h.ncu-rep --page source --csv > $O/tb_half${h}_source.csv 2>/dev/null
done
cat $O/tb_half1_summary.txt
