#!/bin/bash
# sentinel hand-off rows in the lone-pair fill: tests, same-box A/B
O=gpurun_out/sent
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "SENT or golden or band or special or empty or strip_and_macro or fuzz" 2>&1 | tail -5 | tee $O/pytest.txt
export SEQALK_OVERLAP=0
timeout 300 python scripts/ab_env.py cfg5 4 "SEQALK_FILL16_SENT=0" "" "SEQALK_FILL16_SENT=0" "" 2>&1 | tee $O/ab_cfg5.txt
