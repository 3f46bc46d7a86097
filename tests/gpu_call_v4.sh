#!/bin/bash
set -x
V=${1:-r2_v4}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_band_solver.py tests/test_gpu_ba.py tests/test_gpu_full_size.py tests/test_gpu_resident_map.py tests/test_gpu_icp.py tests/test_flann_pin.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > $O/pytest_sel_$V.txt; cat $O/pytest_sel_$V.txt
LVB_KNN_VARIANT=1 timeout 200 python bench.py --skip-global --skip-roofline --skip-cpu > $O/bench_knn1_$V.json 2>/dev/null; python - <<'PY'
import json,glob
d=json.loads(open(glob.glob('gpurun_out/bench_knn1_*.json')[-1]).read().strip().splitlines()[-1]); print('knn rings:', d['icp']['ms_per_scan'], d['kernels']['icp_us_per_scan'])
PY
timeout 400 python bench.py > $O/bench_$V.json 2> $O/bench_$V.err; tail -c 600 $O/bench_$V.json; tail -5 $O/bench_$V.err
python - <<'PY'
import json,glob
d=json.loads(open(sorted(glob.glob('gpurun_out/bench_r2_v4*.json'))[-1]).read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('global', d['global_ba']['ms_per_iteration'], d['global_ba']['cost'], d['kernels']['global_ba_us_per_iteration'])
print('icp', d['icp']['ms_per_scan'], d['icp']['e2e']['ms_per_scan'], d['kernels']['icp_us_per_scan'])
print('w20', d['window20']['ms_per_step'], d['window20']['vs_cpu'])
PY
