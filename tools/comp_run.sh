#!/bin/bash
# tools/comp_run.sh — device session for the compositor: parity tests, C4 lines (both backgrounds), ncu counters; then bench.py
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_comp_gpu.py tests/test_host_paths_gpu.py tests/test_abi.py -q -x -p no:cacheprovider -n 4 2>&1 | tail -3
timeout 300 python bench_extra.py --only c4 --no-cpu 2>&1 | tail -2 | cut -c1-420 | tee $O/c4_prefetch.json
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:comp_kernel -s 3 -c 1 --csv --log-file $O/comp_prefetch_metrics.csv python bench_extra.py --only c4 --no-cpu --steps 3 > $O/comp_prefetch_ncu.log 2>&1; echo "ncu rc=$?"
python tools/lab_metrics.py $O/comp_prefetch_metrics.csv 2>/dev/null | head -4
timeout 900 python bench.py > $O/bench_c2_v2.json 2> $O/bench_c2_v2.err; echo "bench rc=$?"; cut -c1-1500 $O/bench_c2_v2.json
