#!/bin/bash
# Runs on the GPU box: fill-kernel variant sweep on one workload (GCUPS per variant).
WL=${1:-cfg3}
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for R in ${SWEEP_R:-8 16}; do for TAB in ${SWEEP_TAB:-1 0}; do for ADD in ${SWEEP_ADD:-0 1}; do
  export SEQALK_FILL_R=$R SEQALK_FILL_TAB=$TAB SEQALK_FILL_ADD=$ADD
  python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print('R=%s TAB=%s ADD=%s value=%.0f e2e=%.0f fill_gcups=%.0f fill_ms=%.1f tb_ms=%.1f dg_ms=%.1f step_ms=%.1f' % (os.environ['SEQALK_FILL_R'],os.environ['SEQALK_FILL_TAB'],os.environ['SEQALK_FILL_ADD'],d['value'],d['e2e']['value'],d['roofline']['fill_gcups_per_gpu'],d['kernel_ms_per_step']['fill'],d['kernel_ms_per_step']['traceback'],d['kernel_ms_per_step']['digest'],d['ms_per_step']))" | tee -a gpurun_out/sweep.txt
done; done; done
