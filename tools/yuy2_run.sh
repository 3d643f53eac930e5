#!/bin/bash
# packed 4:2:2 -> 4:2:0: parity, then timing
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_vcs_rgbin_gpu.py -q -x -m gpu -n 6 2>&1 | tail -3 | tee $O/yuy2_tests.txt
timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -2 | cut -c1-330 | tee $O/yuy2_bench.json
