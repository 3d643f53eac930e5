#!/bin/bash
# tools/lab_run.sh — one device session of the headline-kernel lab: timings, per-variant ncu counters, one full capture
# with source correlation.  Everything lands in gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $O/lab_clocks.csv &
SMI=$!
timeout 300 tools/l2lab "${1:-}" 20 > $O/l2lab.txt 2>&1; echo "lab rc=$?"; cat $O/l2lab.txt
kill $SMI
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:vcs_ --csv --log-file $O/l2lab_metrics.csv tools/l2lab "${1:-}" 1 > $O/l2lab_ncu.log 2>&1; echo "ncu metrics rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vcs_ -c 1 -f -o $O/l2_full tools/l2lab ref 1 > $O/l2_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -8
