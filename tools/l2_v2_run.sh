#!/bin/bash
# tools/l2_v2_run.sh — device session for the second form of the headline kernel: parity tests, lab timings, ncu counters.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_vcs_lanczos2_gpu.py tests/test_vcs_gpu.py -q -x -p no:cacheprovider -n 4 2>&1 | tail -3
timeout 300 tools/l2lab "${1:-}" 20 > $O/l2lab_v2.txt 2>&1; echo "lab rc=$?"; cat $O/l2lab_v2.txt
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:vcs_lanczos2 --csv --log-file $O/l2lab_v2_metrics.csv tools/l2lab v2_ 1 > $O/l2lab_v2_ncu.log 2>&1; echo "ncu metrics rc=$?"
python tools/lab_metrics.py $O/l2lab_v2_metrics.csv 2>/dev/null | head -12
