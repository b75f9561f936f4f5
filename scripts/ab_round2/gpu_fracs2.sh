#!/bin/bash
O=gpurun_out/${1:-fracs2}
mkdir -p $O
run() { local label="$1"; shift; env "$@" python scripts/shard_span.py cfg4 8 14 3 2>&1 | tail -2 | sed "s/^/$label /" | sed 's/overlap=default cfg4 rank 0 of 8, //; s/(kernel events[^)]*)//' | cut -c1-300 | tee -a $O/span.txt; }
run "cur14    "
run "geo14 .80" SHARD_FRACS=100.00,80.00,64.00,51.20,40.96,32.77,26.21,20.97,16.78,13.42,10.74,8.59,6.87,5.50
run "geo12 .80" SHARD_FRACS=100.00,80.00,64.00,51.20,40.96,32.77,26.21,20.97,16.78,13.42,10.74,8.59
run "geo16 .80" SHARD_FRACS=100.00,80.00,64.00,51.20,40.96,32.77,26.21,20.97,16.78,13.42,10.74,8.59,6.87,5.50,4.40,3.52
run "geo14 .85" SHARD_FRACS=100.00,85.00,72.25,61.41,52.20,44.37,37.71,32.06,27.25,23.16,19.69,16.73,14.22,12.09
run "geo12 .75" SHARD_FRACS=100,75,56.25,42.19,31.64,23.73,17.8,13.35,10.01,7.51,5.63,4.22
