#!/bin/bash
# strip queue order A/B: (strip, pair) vs remaining-strips-first
O=gpurun_out/${1:-queue}
mkdir -p $O
for w in cfg2 cfg3; do
  for q in 0 1; do
    SEQALK_QUEUE_REMAINING=$q SEQALK_OVERLAP=0 python scripts/ab_lib.py $w 3 2>&1 | sed "s/^/remaining=$q /" | tee -a $O/ab.txt
  done
done
for q in 0 1; do
  SEQALK_QUEUE_REMAINING=$q python scripts/ab_lib.py cfg4 2 2>&1 | sed "s/^/remaining=$q overlap /" | tee -a $O/ab.txt
done
