#!/bin/bash
set -x
V=${1:-v4}
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > $O/bench_r1_$V.json 2> $O/bench_r1_$V.err; tail -c 600 $O/bench_r1_$V.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_r1_$V.csv python bench.py --steps 4 --warmup 3 --skip-global > $O/launches_bench_$V.log 2>&1
LVB_NO_GRAPH=1 ncu --set full --clock-control none -k regex:'ba_(linearize|schur_kernel|cholesky|update|build_S|post|prepare)' -c 11 -f -o $O/prof_ba_r1_$V python tools/prof_ba.py > $O/prof_ba_$V.log 2>&1
ncu --set full --clock-control none -k regex:'lidar_|voxel_|ror_|ransac_|imu_preintegrate|cloud_' -c 60 -f -o $O/prof_lidar_r1_$V python tools/prof_lidar.py > $O/prof_lidar_$V.log 2>&1
for r in ba lidar; do ncu -i $O/prof_${r}_r1_$V.ncu-rep --page raw --csv > $O/prof_${r}_r1_${V}_raw.csv 2>/dev/null; done
rm -f $O/prof_lidar_r1_$V.ncu-rep $O/prof_ba_r1_$V.ncu-rep
du -sh $O
