#!/bin/bash
# packed RGB planes through the word-wide scaler: parity, then timing beside the byte-wise kernel
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_vcs_rgbin_gpu.py tests/test_vcs_planes_gpu.py -q -x -m gpu 2>&1 | tail -4 | tee $O/planes3_tests.txt
echo "== fast"; timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -3 | cut -c1-330 | tee $O/planes3_fast.json
echo "== byte-wise"; B200_PLANES_SLOW=1 timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -3 | cut -c1-330 | tee $O/planes3_slow.json
