#!/bin/bash
# A/B of the half-warp tile traceback on one B200 (same box, same process set-up)
O=gpurun_out/${1:-tbh}
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
for h in 0 1; do
  SEQALK_TB_HALF=$h SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg3 2 2>&1 | sed "s/^/half=$h /" | tee -a $O/ab.txt
done
for h in 0 1; do
  SEQALK_TB_HALF=$h SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg4 1 2>&1 | sed "s/^/half=$h /" | tee -a $O/ab.txt
done
