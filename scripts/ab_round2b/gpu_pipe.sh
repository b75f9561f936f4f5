#!/bin/bash
# packed fill for few long pairs: predicated checkpoint stores / pipelined hand-off / resident CTAs: tests, same-box A/B, ncu of the lone pair
O=gpurun_out/pipe
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "MINB" 2>&1 | tail -5 | tee $O/pytest.txt
export SEQALK_OVERLAP=0
timeout 300 python scripts/ab_env.py cfg5 3 "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=28" "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=38" "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=34" \
  "SEQALK_FILL_MINB=36" "SEQALK_FILL_MINB=33" "SEQALK_FILL16_RP=4" 2>&1 | tee $O/ab_cfg5.txt
timeout 300 python scripts/ab_env.py cfg2 3 "" "SEQALK_FILL_MINB=23" "SEQALK_FILL_MINB=33" "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=28" "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=38" \
  "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=34" "SEQALK_FILL_MINB=23,SEQALK_FILL_CTAS_PER_SM=2" "SEQALK_FILL_MINB=23,SEQALK_FILL_CTAS_PER_SM=3" \
  "SEQALK_FILL_MINB=33,SEQALK_FILL_CTAS_PER_SM=2" 2>&1 | tee $O/ab_cfg2.txt
unset SEQALK_OVERLAP
timeout 400 python scripts/ab_env.py cfg4 2 "" "SEQALK_FILL_MINB=23" "SEQALK_FILL_MINB=33" 2>&1 | tee $O/ab_cfg4.txt
export SEQALK_OVERLAP=0
SEQALK_FILL16_RP=8 SEQALK_FILL_MINB=28 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fill16_all -c 1 -o $O/fill16_cfg5_rp8 -f \
   python scripts/ab_env.py cfg5 1 "" > $O/ncu_full.log 2>&1
ls -la $O
