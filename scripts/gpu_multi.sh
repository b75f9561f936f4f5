#!/bin/bash
# Runs on a multi-GPU box: multi-GPU parity tests, then the torchrun bench at N ranks.
# usage: bash scripts/gpu_multi.sh <N> <workload> [steps] [warmup]
N=${1:-2}; WL=${2:-cfg3}; K=${3:-2}; W=${4:-3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or two_ranks or distributed" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --workload $WL --steps $K --warmup $W > gpurun_out/bench_${WL}_n$N.json 2> gpurun_out/bench_${WL}_n$N.err
echo "bench rc=$?"; grep -B2 -A12 "Traceback" gpurun_out/bench_${WL}_n$N.err | grep -v "^W10" | head -40
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${WL}_n$N.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ["value","n_gpus","ms_per_step","gpu_launches"]}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "fill/gpu", d["roofline"]["fill_gcups_per_gpu"], d["kernel_ms_per_step"], d["digest_e2e_matches"], d["digest"][:16])
except Exception as e: print("no json", e)
PY
