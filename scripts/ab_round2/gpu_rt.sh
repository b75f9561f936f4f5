#!/bin/bash
O=gpurun_out/${1:-rt}
mkdir -p $O
SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg3 2 2>&1 | tee -a $O/ab.txt
SEQALK_OVERLAP=0 python scripts/ab_lib.py cfg4 1 2>&1 | tee -a $O/ab.txt
python scripts/ab_lib.py cfg4 2 2>&1 | sed "s/^/overlap /" | tee -a $O/ab.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
