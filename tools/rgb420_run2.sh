#!/bin/bash
# packed RGB -> 4:2:0 (all three forms) and the 4-byte plane scaler with straight-line taps: parity, then timing
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_vcs_rgbin_gpu.py -q -x -m gpu -n 4 2>&1 | tail -3 | tee $O/rgb420b_tests.txt
timeout 300 python tools/gpu_cross_check.py 60 rgb 2>&1 | tail -3 | tee $O/rgb420b_cross.txt
echo "== fast"; timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -9 | cut -c1-330 | tee $O/rgb420b_fast.json
