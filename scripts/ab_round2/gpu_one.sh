#!/bin/bash
# one-launch packed fill vs one launch per strip height; 128 x 128 tiles; same box
O=gpurun_out/${1:-one}
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
for sp in 1 0; do
  SEQALK_FILL16_SPLIT=$sp SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg3 2 2>&1 | sed "s/^/split=$sp /" | tee -a $O/ab.txt
  SEQALK_FILL16_SPLIT=$sp SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg4 1 2>&1 | sed "s/^/split=$sp /" | tee -a $O/ab.txt
  SEQALK_FILL16_SPLIT=$sp python scripts/ab_lib.py cfg4 2 2>&1 | sed "s/^/split=$sp overlap /" | tee -a $O/ab.txt
  SEQALK_FILL16_SPLIT=$sp python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/split=$sp /" | tee -a $O/span.txt
done
export SEQALK_LIB=$PWD/ab/ckq32/libseqalk.so
for h in 0 1; do
  SEQALK_TB_HALF=$h SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg3 2 2>&1 | sed "s/^/ckq32 half=$h /" | tee -a $O/ab.txt
  SEQALK_TB_HALF=$h SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg4 1 2>&1 | sed "s/^/ckq32 half=$h /" | tee -a $O/ab.txt
done
