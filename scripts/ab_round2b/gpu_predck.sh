#!/bin/bash
# predicated checkpoint stores in the packed fill: parity tests, same-box A/B on cfg5 / cfg2 / cfg4
O=gpurun_out/predck
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "MINB or band or cfg5_full_size_golden or golden_big13" 2>&1 | tail -8 | tee $O/pytest.txt
export SEQALK_OVERLAP=0
timeout 300 python scripts/ab_env.py cfg5 3 "" "SEQALK_FILL_MINB=26" "SEQALK_FILL_MINB=23" "SEQALK_FILL16_RP=8,SEQALK_FILL_MINB=28" "SEQALK_FILL_MINB=3" 2>&1 | tee $O/ab_cfg5.txt
timeout 300 python scripts/ab_env.py cfg2 3 "" "SEQALK_FILL_MINB=26" "SEQALK_FILL_MINB=23" "SEQALK_FILL_MINB=3" "SEQALK_FILL_MINB=4" "SEQALK_FILL16_RP=32,SEQALK_FILL_MINB=24" 2>&1 | tee $O/ab_cfg2.txt
timeout 400 python scripts/ab_env.py cfg4 2 "" "SEQALK_FILL_MINB=24" "SEQALK_FILL_MINB=23" 2>&1 | tee $O/ab_cfg4.txt
