#!/bin/bash
set -x
V=${1:-r2_v5}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_resident_map.py tests/test_gpu_icp.py tests/test_flann_pin.py tests/test_gpu_shim.py tests/test_zy_ref_golden_cuda.py tests/test_zz_ref_backend_dropin.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > $O/pytest_sel_$V.txt; cat $O/pytest_sel_$V.txt
for kv in 0 1; do
LVB_KNN_VARIANT=$kv timeout 200 python bench.py --skip-global --skip-roofline --skip-cpu > $O/bench_knn${kv}_$V.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_knn${kv}_$V.json').read().strip().splitlines()[-1]); print('knn variant $kv:', d['icp']['ms_per_scan'], d['icp']['e2e']['ms_per_scan'], d['kernels']['icp_us_per_scan'])
PY
done
