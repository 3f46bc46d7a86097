for cfg in "LVB_NO_SIDE=1" "LVB_NO_SIDE=1 LVB_FUSE_PREPARE=1" "LVB_NO_SIDE=0" "LVB_NO_SIDE=0 LVB_NO_GRAPH=1" "LVB_NO_SIDE=1 LVB_NO_GRAPH=1"; do
  echo "== $cfg"; env $cfg timeout 200 python bench.py --skip-global --skip-roofline --skip-icp 2>&1 | grep -o '"value": [0-9.]*, "unit": "rows/s", "n_gpus": 1, "steps": 40, "warmup": 10, "ms_per_step": [0-9.]*'
done
