#!/bin/bash
set -x
V=${1:-r2_v10}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_band_solver.py tests/test_gpu_ba.py tests/test_gpu_full_size.py tests/test_gpu_tensor_schur.py tests/test_gpu_shim.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > $O/pytest_sel_$V.txt; cat $O/pytest_sel_$V.txt
python tools/chol_clocks.py
timeout 400 python bench.py --skip-icp --skip-cpu --skip-roofline > $O/bench_$V.json 2> $O/bench_$V.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$V.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('w20', d['window20']['ms_per_step'], d['window20']['e2e']['value'])
print('global', d['global_ba']['ms_per_iteration'], d['kernels']['global_ba_us_per_iteration'])
print(d['kernels']['window10_us_per_iteration'])
PY
