#!/bin/bash
for v in 0 1 2 3; do
  LVB_EVAL_VARIANT=$v python bench.py --steps 4 --warmup 3 --skip-icp 2>&1 | tail -1 > /tmp/ab_$v.json
  python -c "import json; d=json.load(open('/tmp/ab_$v.json')); print('variant', $v, d['roofline']['us_per_launch'], d['roofline']['frac'])"
done
python -m pytest tests/test_gpu_ba.py -m gpu -q -x 2>&1 | tail -2
