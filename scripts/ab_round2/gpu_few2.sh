#!/bin/bash
O=gpurun_out/${1:-few2}
mkdir -p $O
for w in cfg2 cfg5; do
  SEQALK_OVERLAP=0 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/base /" | tee -a $O/ab.txt
  SEQALK_OVERLAP=0 SEQALK_FILL_R=32 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/fill16 rp32 one-launch /" | tee -a $O/ab.txt
  SEQALK_OVERLAP=0 SEQALK_FILL_R=32 SEQALK_FILL16_RP=16 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/fill16 rp16 one-launch /" | tee -a $O/ab.txt
done
