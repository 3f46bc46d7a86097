#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU):  bash tools/collect_profiles_r2.sh r2_v1
#   launch list of the bench command (per-launch device time, cold cache, serialised: compare SHARES) and --set full captures of
#   the kernels the round changed: the multifrontal Cholesky, the fused linearisation at map scale, the window pass, the ICP tile kernel.
set -x
V=${1:-r2_v1}
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_$V.csv python bench.py --steps 4 --warmup 3 --blocks 1 --skip-cpu > $O/launches_bench_$V.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ba_(front_|linearize_kernel|schur_kernel|update|build_S)' -s 0 -c 40 -f -o $O/prof_global_$V python tools/prof_global.py > $O/prof_global_$V.log 2>&1
LVB_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:'ba_(linearize|schur_kernel|cholesky|update|build_S|post)' -c 10 -f -o $O/prof_ba_$V python tools/prof_ba.py > $O/prof_ba_$V.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'icp_(tile|associate|linearize)' -c 4 -f -o $O/prof_icp_$V python tools/prof_icp.py > $O/prof_icp_$V.log 2>&1
for r in global ba icp; do ncu -i $O/prof_${r}_$V.ncu-rep --page raw --csv > $O/prof_${r}_${V}_raw.csv 2>/dev/null; done
# the merge-back limit is 64 MiB: keep the raw CSVs (and the per-instruction source page of the two kernels of interest), drop the reports
ncu -i $O/prof_ba_$V.ncu-rep --page source --csv -k regex:ba_cholesky > $O/prof_ba_${V}_cholesky_source.csv 2>/dev/null
ncu -i $O/prof_global_$V.ncu-rep --page source --csv -k regex:ba_linearize_kernel -c 1 > $O/prof_global_${V}_linearize_source.csv 2>/dev/null
rm -f $O/prof_global_$V.ncu-rep $O/prof_ba_$V.ncu-rep $O/prof_icp_$V.ncu-rep
ls -la $O | tail -20
du -sh $O
