#!/bin/bash
O=gpurun_out/${1:-cfg2val}
mkdir -p $O
run() { local label="$1"; shift; env "$@" python bench.py --workload cfg2 --steps 4 --warmup 2 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$label', 'value ms', round(d['ms_per_step'],1), 'e2e ms', round(d['e2e']['ms_per_step'],1), d['kernel_ms_per_step']['fill'], d['e2e'].get('phases_ms'))" | tee -a $O/out.txt; }
run "default"
run "slots2" SEQALK_SLOTS=2
run "prio0" SEQALK_POST_PRIORITY=0
run "slots2 prio0" SEQALK_SLOTS=2 SEQALK_POST_PRIORITY=0
run "fill16" SEQALK_FILL_R=32 SEQALK_FILL16_RP=16
