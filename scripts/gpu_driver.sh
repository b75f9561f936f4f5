#!/bin/bash
# Runs on a GPU box: the C++ stdin/stdout driver (seqalkway) on a named workload with 1..N GPUs
# of one process.  usage: bash scripts/gpu_driver.sh <workload> <N> [N...]
WL=${1:-cfg4}; shift
mkdir -p gpurun_out
python - <<PY
import importlib
w = importlib.import_module("multiple-sequence-alignment-openmp-openmpi_b200.workloads")
pxy, pgap, genes = w.make("$WL")
open("/tmp/$WL.txt", "wb").write(w.format_input(pxy, pgap, genes))
PY
for n in "$@"; do
  echo "== $WL, $n GPU(s) in one process"
  SEQALK_GPUS=$n timeout 600 ./multiple-sequence-alignment-openmp-openmpi_b200/seqalkway < /tmp/$WL.txt > gpurun_out/driver_${WL}_n$n.out 2> gpurun_out/driver_${WL}_n$n.err
  echo "rc=$?"; head -c 300 gpurun_out/driver_${WL}_n$n.out | head -2; tail -c 200 gpurun_out/driver_${WL}_n$n.out | tail -1 | cut -c1-80
done
