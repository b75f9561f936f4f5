#!/bin/bash
O=gpurun_out/${1:-ckq2}
mkdir -p $O
for w in cfg2 cfg5; do
  SEQALK_OVERLAP=0 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/ckq64 /" | tee -a $O/ab.txt
  SEQALK_LIB=$PWD/ab/ckq32/libseqalk.so SEQALK_OVERLAP=0 python scripts/ab_lib.py $w 2 2>&1 | sed "s/^/ckq32 /" | tee -a $O/ab.txt
done
SEQALK_LIB=$PWD/ab/ckq32/libseqalk.so SEQALK_TB_HALF=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/pytest_ckq32_half.txt; cat $O/pytest_ckq32_half.txt
