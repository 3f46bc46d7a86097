#!/bin/bash
# The first gpurun call of the next round: everything that was only cross-compiled at the end of
# round 1 runs here, late-sorting files last so that a failure there cannot hide the verified tests.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tests/first_gpu_call.sh r2_v0'
set -x
V=${1:-r2_v0}
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > $O/pytest_gpu_$V.txt; cat $O/pytest_gpu_$V.txt
# the late files on their own as well, without -x, so that every one of the new tests reports
timeout 240 python -m pytest tests/test_zy_ref_golden_cuda.py tests/test_zz_ref_backend_dropin.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > $O/pytest_new_$V.txt; cat $O/pytest_new_$V.txt
for m in 0 1 2 3; do timeout 60 oracle/_ref/ref_backend_lvb $m > $O/ref_backend_lvb_$m.txt 2>&1; timeout 60 oracle/_ref/ref_backend_orc $m > $O/ref_backend_orc_$m.txt 2>&1; done
timeout 60 oracle/_ref/ref_mapping_lvb > $O/ref_mapping_lvb.txt 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > $O/bench_$V.json 2> $O/bench_$V.err; tail -c 800 $O/bench_$V.json
du -sh $O
