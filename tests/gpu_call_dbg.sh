#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/pytest_dbg.txt 2>&1
grep -n "Fatal\|Segmentation\|test_\|Error" $O/pytest_dbg.txt | head -40
