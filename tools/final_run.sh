#!/bin/bash
# last device call of the round: the whole device suite, the bench line, one full capture of the newest hot kernel
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -q -m gpu -n 6 -p no:cacheprovider 2>&1 | tail -6 > $O/final2_gputests.txt; tail -2 $O/final2_gputests.txt
timeout 170 python bench.py > $O/final2_bench_c2.json 2> $O/final2_bench_c2.err; cut -c1-400 $O/final2_bench_c2.json
timeout 100 ncu --set full --clock-control none -c 1 -f -k regex:vcs_rgb420 -o $O/final2_rgb420 python bench_extra.py --only planes --no-cpu --steps 1 > $O/final2_rgb420.log 2>&1
python tools/ncu_summary.py $O/final2_rgb420.ncu-rep $O/final2_rgb420_ncu.txt > /dev/null 2>&1 && rm -f $O/final2_rgb420.ncu-rep
ls -la $O | grep final2
