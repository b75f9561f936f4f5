#!/bin/bash
# Runs on the GPU box: pipeline knob sweep (slots / headroom / budget) on one workload.
WL=${1:-cfg3}
mkdir -p gpurun_out
run() {
  env "$@" python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', 'value=%.0f e2e=%.0f fill_gcups=%.0f fill_ms=%.1f tb_ms=%.1f dg_ms=%.1f step_ms=%.1f' % (d['value'],d['e2e']['value'],d['roofline']['fill_gcups_per_gpu'],d['kernel_ms_per_step']['fill'],d['kernel_ms_per_step']['traceback'],d['kernel_ms_per_step']['digest'],d['ms_per_step']))" | tee -a gpurun_out/knobs.txt
}
: > gpurun_out/knobs.txt
run SEQALK_SLOTS=2
run SEQALK_SLOTS=3
run SEQALK_SLOTS=3 SEQALK_DIR_BUDGET_GB=40
run SEQALK_SLOTS=2 SEQALK_FILL_HEADROOM_REGS=12288
run SEQALK_SLOTS=3 SEQALK_FILL_HEADROOM_REGS=12288
run SEQALK_SLOTS=2 SEQALK_FILL_R=8
run SEQALK_SLOTS=3 SEQALK_FILL_R=8
