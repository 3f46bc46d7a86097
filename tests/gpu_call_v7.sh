#!/bin/bash
set -x
V=${1:-r2_v7}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > $O/pytest_gpu_$V.txt; cat $O/pytest_gpu_$V.txt
python tools/chol_clocks.py
timeout 200 python bench.py --quick > $O/bench_quick_$V.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_quick_$V.json').read().strip().splitlines()[-1]); print('quick:', d['ms_per_step'], d['value'], d['e2e']['value'], d['kernels'])
PY
LVB_NO_MERGED_LINEARIZE=1 timeout 200 python bench.py --quick > $O/bench_quick_nomerge_$V.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_quick_nomerge_$V.json').read().strip().splitlines()[-1]); print('quick no-merge:', d['ms_per_step'], d['value'], d['e2e']['value'])
PY
timeout 400 python bench.py > $O/bench_$V.json 2> $O/bench_$V.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$V.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print('w20', d['window20']['ms_per_step'], d['window20']['vs_cpu'])
print('global', d['global_ba']['ms_per_iteration'])
print('icp', d['icp']['ms_per_scan'], d['icp']['e2e']['ms_per_scan'], d['icp']['vs_cpu'])
PY
