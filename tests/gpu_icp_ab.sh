for sh in 2 1 0; do
LVB_ICP_SORT_SHIFT=$sh timeout 200 python bench.py --skip-global --skip-roofline --skip-cpu > gpurun_out/bench_icp_shift$sh.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_icp_shift$sh.json').read().strip().splitlines()[-1]); print('shift $sh:', d['icp']['ms_per_scan'], d['icp']['e2e_resident_map']['ms_per_keyframe'], d['kernels']['icp_us_per_scan'])
PY
done
timeout 300 python -m pytest tests/test_gpu_icp.py tests/test_gpu_resident_map.py tests/test_flann_pin.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
