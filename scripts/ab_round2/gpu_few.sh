#!/bin/bash
# few-pair configs (cfg2 = big13, cfg5 = one 200k pair): 32-bit fill vs packed fill, same box
O=gpurun_out/${1:-few}
mkdir -p $O
for w in cfg2 cfg5; do
  SEQALK_OVERLAP=0 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/base /" | tee -a $O/ab.txt
  SEQALK_OVERLAP=0 SEQALK_FILL_R=32 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/R32(fill16 rp32) /" | tee -a $O/ab.txt
  SEQALK_OVERLAP=0 SEQALK_FILL_R=32 SEQALK_FILL16_RP=16 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/R32(fill16 rp16) /" | tee -a $O/ab.txt
  SEQALK_OVERLAP=0 SEQALK_FILL_R=8 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/R8(32-bit) /" | tee -a $O/ab.txt
done
