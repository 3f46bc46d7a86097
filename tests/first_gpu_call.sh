#!/bin/bash
# One-GPU check of everything (the driver's round-end sequence): pytest -m gpu, smoke, the default bench.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tests/first_gpu_call.sh r2_v3'
set -x
V=${1:-r2_v0}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 2>&1 | tail -40 > $O/pytest_gpu_$V.txt; cat $O/pytest_gpu_$V.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > $O/bench_$V.json 2> $O/bench_$V.err; tail -c 1500 $O/bench_$V.json; tail -5 $O/bench_$V.err
du -sh $O
