#!/bin/bash
# 8-GPU box: the torchrun bench line, then the in-process call with phase marks (default chunks)
O=gpurun_out/${1:-n8}
N=${2:-8}
mkdir -p $O
nproc > $O/nproc.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 2 --warmup 2 > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench rc=$?"; head -c 300 $O/bench_n$N.json
SEQALK_TRACE=1 python scripts/peer_chunks.py cfg4 $N ${3:-0} 2 > $O/peer_chunks.txt 2> $O/peer_chunks.err
cat $O/peer_chunks.txt
