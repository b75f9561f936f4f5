#!/bin/bash
# sentinel hand-off rows also for 1024-row strip pairs (big13): tests, same-box A/B
O=gpurun_out/sent2
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "variant or golden or band or special or empty or strip_and_macro or fuzz or batching or tiny" 2>&1 | tail -5 | tee $O/pytest.txt
export SEQALK_OVERLAP=0
timeout 300 python scripts/ab_env.py cfg2 3 "SEQALK_FILL16_SENT=0" "" "SEQALK_FILL_CTAS_PER_SM=3" "SEQALK_FILL_CTAS_PER_SM=4" "SEQALK_FILL16_SENT=0" "" 2>&1 | tee $O/ab_cfg2.txt
timeout 300 python scripts/ab_env.py cfg5 3 "" 2>&1 | tee $O/ab_cfg5.txt
