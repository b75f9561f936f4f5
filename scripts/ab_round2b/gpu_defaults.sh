#!/bin/bash
# new defaults (fill builds per job shape, banded traceback, host hashing): whole GPU suite, then same-box A/B against the old path
O=gpurun_out/defaults
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest.txt
export SEQALK_OVERLAP=0
OLD="SEQALK_TB_BAND=0,SEQALK_HOST_HASH=0,SEQALK_FILL16_RP=16,SEQALK_FILL16_PREDCK=0,SEQALK_FILL_CTAS_PER_SM=0"
timeout 300 python scripts/ab_env.py cfg5 3 "$OLD" "" "SEQALK_FILL_MINB=8" 2>&1 | tee $O/ab_cfg5.txt
timeout 300 python scripts/ab_env.py cfg2 3 "$OLD" "" "SEQALK_FILL_CTAS_PER_SM=1" "SEQALK_FILL_CTAS_PER_SM=3" "SEQALK_FILL16_RP=8,SEQALK_FILL_CTAS_PER_SM=3" "SEQALK_FILL16_RP=8,SEQALK_FILL_CTAS_PER_SM=2" 2>&1 | tee $O/ab_cfg2.txt
timeout 300 python scripts/ab_env.py cfg3 3 "SEQALK_FILL16_PREDCK=0" "" "SEQALK_FILL_MINB=4" 2>&1 | tee $O/ab_cfg3.txt
for w in cfg5 cfg2; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$w.csv \
   python scripts/ab_env.py $w 1 "" > $O/ncu_$w.log 2>&1
done
