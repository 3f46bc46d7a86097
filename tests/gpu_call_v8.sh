#!/bin/bash
set -x
V=${1:-r2_v8}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > $O/pytest_gpu_$V.txt; cat $O/pytest_gpu_$V.txt
LVB_PROFILE=1 timeout 100 python - <<'PY' 2>&1 | tail -30
import sys; sys.path.insert(0, '.')
from lvio_fusion_b200 import _capi, backend, synth
import time
lvb=_capi.load(); ctx=backend.Context(lvb)
d=synth.make_ba_problem(10,4000,with_imu=True)
for i in range(3):
    t0=time.perf_counter(); p=backend.Problem.from_dict(ctx,d); t1=time.perf_counter(); s=p.solve(max_num_iterations=10,function_tolerance=0.0,gradient_tolerance=0.0,parameter_tolerance=0.0); t2=time.perf_counter(); p.poses(); p.vec3(); p.inv_depths(); t3=time.perf_counter(); p.close()
    print("from_dict %.0f us  solve %.0f us  download %.0f us"%((t1-t0)*1e6,(t2-t1)*1e6,(t3-t2)*1e6), file=sys.stderr)
PY
timeout 400 python bench.py > $O/bench_$V.json 2> $O/bench_$V.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$V.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print('w20', d['window20']['ms_per_step'], d['window20']['e2e']['value'], d['window20']['vs_cpu'])
print('global', d['global_ba']['ms_per_iteration'], d['kernels']['global_ba_us_per_iteration'])
print('fused', d.get('roofline_fused'))
print('tc', d.get('schur_tc'))
print('icp', d['icp']['ms_per_scan'], d['icp']['e2e']['ms_per_scan'], d['icp']['vs_cpu'])
PY
tail -3 $O/bench_$V.err
LVB_NO_GRAPH_CACHE=1 timeout 200 python bench.py --quick > $O/bench_quick_nocache_$V.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_quick_nocache_$V.json').read().strip().splitlines()[-1]); print('no graph cache:', d['ms_per_step'], d['value'], d['e2e']['value'])
PY
