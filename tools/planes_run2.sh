#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
echo "== fast"; timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | cut -c1-330 | tee $O/planes_fast.json
echo "== generic chain for the cross-family pairs"; B200_CROSS_GENERIC=1 timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | tail -2 | cut -c1-330 | tee $O/cross_generic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/planes_launches.csv python bench_extra.py --only planes --no-cpu --steps 1 > $O/planes_launches.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $O/planes_launches.csv | cut -d, -f5,9,15 | sort | uniq -c | sort -rn | head -30
