#!/bin/bash
# packed RGB -> 4:2:0 fast kernels: parity, then timing beside the generic chain
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_vcs_rgbin_gpu.py -q -x -m gpu 2>&1 | tail -4 | tee $O/rgb420_tests.txt
timeout 600 python tools/gpu_cross_check.py 75 rgb 2>&1 | tail -3 | tee $O/rgb420_cross.txt
echo "== fast"; timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -5 | cut -c1-330 | tee $O/rgb420_fast.json
echo "== generic"; B200_RGB420_GENERIC=1 timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -5 | cut -c1-330 | tee $O/rgb420_generic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/rgb420_launches.csv python bench_extra.py --only planes --no-cpu --steps 1 > $O/rgb420_launches.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $O/rgb420_launches.csv | grep -i "rgb420\|planes_fast" | cut -d, -f5,9,15 | sort | uniq -c | sort -rn | head -30 | tee $O/rgb420_launch_summary.txt
rm -f $O/rgb420_launches.csv
