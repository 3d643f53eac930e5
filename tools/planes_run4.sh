#!/bin/bash
# h-first plane scaler with vector T accesses and straight-line taps: parity of everything that runs it, then timing
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_vcs_rgbin_gpu.py tests/test_vcs_planes_gpu.py tests/test_vcs_cross_gpu.py -q -x -m gpu -n 6 2>&1 | tail -3 | tee $O/planes4_tests.txt
echo "== fast"; timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -15 | cut -c1-330 | tee $O/planes4_fast.json
