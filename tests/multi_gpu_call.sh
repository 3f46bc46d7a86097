#!/bin/bash
# The multi-GPU check, with the driver's exact bench command (no skip flags):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tests/multi_gpu_call.sh 2 r2_v1'            # bench + tests
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 600 -- 'bash tests/multi_gpu_call.sh 8 r2_v1 bench'      # bench only
set -x
N=${1:-2}
V=${2:-r2_v0}
ONLY=${3:-all}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_$V.json 2> $O/bench_n${N}_$V.err
echo "bench rc=$?"; tail -c 600 $O/bench_n${N}_$V.json; tail -5 $O/bench_n${N}_$V.err
[ "$ONLY" = "bench" ] && exit 0
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider 2>&1 | tail -6 > $O/pytest_multi_n${N}_$V.txt; cat $O/pytest_multi_n${N}_$V.txt
# a collective issued by one rank only must come back as LVB_ERR_COMM (bounded wait), not as a trap
LVB_P2P_TIMEOUT_MS=1500 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tests/multigpu_worker.py 2>&1 | grep -E "MULTIGPU|MISMATCH" > $O/mismatch_$V.txt; cat $O/mismatch_$V.txt
