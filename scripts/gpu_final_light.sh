#!/bin/bash
# smoke + GPU tests + the default bench line (no ncu): bash scripts/gpu_final_light.sh <tag>
O=gpurun_out/${1:-final_light}
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest.txt; cat $O/pytest.txt
python bench.py --steps 2 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
head -c 300 $O/bench_n1.json
