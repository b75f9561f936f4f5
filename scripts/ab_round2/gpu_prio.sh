#!/bin/bash
O=gpurun_out/${1:-prio}
mkdir -p $O
for pr in 0 1; do
  SEQALK_POST_PRIORITY=$pr SEQALK_SLOTS=2 python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/prio=$pr slots=2 /" | tee -a $O/span.txt
  SEQALK_POST_PRIORITY=$pr SEQALK_SLOTS=3 SEQALK_DIR_BUDGET_GB=40 python scripts/shard_span.py cfg4 8 14 3 2>&1 | sed "s/^/prio=$pr slots=3 budget40 /" | tee -a $O/span.txt
done
for pr in 0 1; do
  SEQALK_POST_PRIORITY=$pr SEQALK_SLOTS=2 python scripts/ab_lib.py cfg4 2 2>&1 | sed "s/^/prio=$pr slots=2 /" | tee -a $O/ab.txt
  SEQALK_POST_PRIORITY=$pr SEQALK_SLOTS=3 SEQALK_DIR_BUDGET_GB=40 python scripts/ab_lib.py cfg4 2 2>&1 | sed "s/^/prio=$pr slots=3 budget40 /" | tee -a $O/ab.txt
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
