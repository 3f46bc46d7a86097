#!/bin/bash
set -x
V=${1:-r2_v6}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_band_solver.py tests/test_gpu_ba.py tests/test_gpu_full_size.py tests/test_gpu_resident_map.py tests/test_gpu_icp.py tests/test_flann_pin.py tests/test_gpu_shim.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > $O/pytest_sel_$V.txt; cat $O/pytest_sel_$V.txt
for cap in 2048 4096 6144 9216; do
LVB_TILE_CAP=$cap timeout 200 python bench.py --skip-global --skip-roofline --skip-cpu > $O/bench_cap${cap}_$V.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cap${cap}_$V.json').read().strip().splitlines()[-1]); print('tile cap $cap:', d['icp']['ms_per_scan'], d['icp']['e2e']['ms_per_scan'], d['kernels']['icp_us_per_scan'])
print('   headline', d['ms_per_step'], d['kernels']['window10_us_per_iteration'].get('ba_cholesky_kernel'), 'w20', d['window20']['ms_per_step'], d['kernels']['window20_us_per_iteration'].get('ba_cholesky_kernel'))
PY
done
timeout 400 python bench.py --skip-icp --skip-cpu > $O/bench_$V.json 2> $O/bench_$V.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$V.json').read().strip().splitlines()[-1]); print('global', d['global_ba']['ms_per_iteration'], d['kernels']['global_ba_us_per_iteration'])
PY
