#!/bin/bash
# ncu --set full of the lone pair's fill (512-row strip pairs, predicated stores, pipelined hand-off)
O=gpurun_out/ncu5
mkdir -p $O
export SEQALK_OVERLAP=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fill16_all -c 1 -o $O/fill16_cfg5_rp8 -f \
   python scripts/ab_env.py cfg5 1 "" > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log; ls -la $O
