#!/bin/bash
O=gpurun_out/${1:-n2}
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 2 --warmup 2 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench rc=$?"; head -c 300 $O/bench_n2.json
