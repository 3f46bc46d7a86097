#!/bin/bash
# run on the GPU box through gpurun; writes into gpurun_out/ (keep the total under 64 MiB)
set -x
V=${1:-v3}
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_r1_$V.csv python bench.py --steps 4 --warmup 3 > $O/launches_bench_$V.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ba_eval_two_frame -c 1 -f -o $O/prof_eval_r1_$V python bench.py --steps 1 --warmup 1 --skip-icp --skip-global > $O/prof_eval_$V.log 2>&1
LVB_NO_GRAPH=1 ncu --set full --clock-control none -k regex:'ba_(linearize|schur|cholesky|update|build_S|post|prepare)' -c 11 -f -o $O/prof_ba_r1_$V python tools/prof_ba.py > $O/prof_ba_$V.log 2>&1
LVB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:ba_schur_tc -c 1 -f -o $O/prof_schur_tc_r1_$V python tools/prof_ba.py > $O/prof_schur_tc_$V.log 2>&1
ncu --set full --clock-control none -k regex:'icp_' -c 30 -f -o $O/prof_icp_r1_$V python tools/prof_icp.py > $O/prof_icp_$V.log 2>&1
for r in eval ba schur_tc icp; do ncu -i $O/prof_${r}_r1_$V.ncu-rep --page raw --csv > $O/prof_${r}_r1_${V}_raw.csv 2>/dev/null; done
du -sh $O; ls -la $O | tail -20
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 60 ]; then rm -f $O/prof_ba_r1_$V.ncu-rep $O/prof_icp_r1_$V.ncu-rep; fi
du -sh $O
