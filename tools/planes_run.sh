#!/bin/bash
# tools/planes_run.sh — device session for the word-wide plane scaler: parity tests, bench lines old / new, ncu counters
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_vcs_planes_gpu.py tests/test_vcs_planar_gpu.py tests/test_vcs_cross_gpu.py tests/test_fuzz_gpu.py tests/test_vcs_borders_gpu.py -q -x -p no:cacheprovider -n 6 2>&1 | tail -3
echo "== fast"; timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | cut -c1-330 | tee $O/planes_fast.json
echo "== byte-wise"; B200_CROSS_GENERIC=1 B200_PLANES_SLOW=1 timeout 300 python bench_extra.py --only planes --no-cpu 2>&1 | cut -c1-330 | tee $O/planes_slow.json
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:vcs_planes -s 6 -c 4 --csv --log-file $O/planes_fast_metrics.csv python bench_extra.py --only planes --no-cpu --steps 2 > $O/planes_fast_ncu.log 2>&1; echo "ncu rc=$?"
python tools/lab_metrics.py $O/planes_fast_metrics.csv 2>/dev/null | head -8
