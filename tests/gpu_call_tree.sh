#!/bin/bash
# tree solver bring-up: its own parity test, the BA tests that use banded storage, then graph-vs-direct A/B of the headline leg
set -x
V=${1:-r2_v2}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_band_solver.py -q -x -p no:cacheprovider 2>&1 | tail -25 > $O/pytest_band_$V.txt; cat $O/pytest_band_$V.txt
timeout 300 python -m pytest tests/test_gpu_ba.py -q -p no:cacheprovider 2>&1 | tail -15 > $O/pytest_ba_$V.txt; cat $O/pytest_ba_$V.txt
timeout 200 python bench.py --quick > $O/bench_quick_graph_$V.json 2> $O/bench_quick_graph_$V.err; tail -c 2500 $O/bench_quick_graph_$V.json; tail -3 $O/bench_quick_graph_$V.err
LVB_NO_GRAPH=1 timeout 200 python bench.py --quick > $O/bench_quick_nograph_$V.json 2> $O/bench_quick_nograph_$V.err; tail -c 700 $O/bench_quick_nograph_$V.json
timeout 400 python bench.py > $O/bench_$V.json 2> $O/bench_$V.err; tail -c 3000 $O/bench_$V.json; tail -5 $O/bench_$V.err
