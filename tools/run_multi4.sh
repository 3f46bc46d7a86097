#!/bin/bash
N=${1:-4}
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_worker.py 2>&1 | grep -a "MULTIGPU" | head -3
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 40 --warmup 10 --skip-roofline 2>&1 | tail -c 2500
