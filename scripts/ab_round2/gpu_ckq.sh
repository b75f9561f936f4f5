#!/bin/bash
# 128 x 128 tiles (library built with -DSEQALK_CKQ=32 in ab/ckq32) vs 128 x 256, warp vs half-warp traceback
O=gpurun_out/${1:-ckq}
mkdir -p $O
for w in cfg3 cfg4; do
  for lib in default ab/ckq32/libseqalk.so; do
    for h in 0 1; do
      if [ $lib = default ]; then unset SEQALK_LIB; else export SEQALK_LIB=$PWD/$lib; fi
      SEQALK_TB_HALF=$h SEQALK_OVERLAP=0 python scripts/ab_lib.py $w 1 2>&1 | sed "s/^/half=$h /" | tee -a $O/ab.txt
    done
  done
done
