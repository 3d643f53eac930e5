#!/bin/bash
# tools/l2_v2_run.sh — device session for the second form of the headline kernel and the light kernel: parity tests, lab
# timings, C1 bench line, ncu counters.
set -u
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_vcs_lanczos2_gpu.py tests/test_vcs_light_gpu.py tests/test_vcs_gpu.py tests/test_fuzz_gpu.py -q -x -p no:cacheprovider -n 6 2>&1 | tail -3
timeout 300 tools/l2lab "${1:-v2_}" 20 > $O/l2lab_v2c.txt 2>&1; echo "lab rc=$?"; cat $O/l2lab_v2c.txt
timeout 300 python bench_extra.py --only c1 --no-cpu 2>&1 | tail -1 | cut -c1-500 | tee $O/c1_pairs.json
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:vcs_lanczos2 --csv --log-file $O/l2lab_v2c_metrics.csv tools/l2lab v2_ 1 > $O/l2lab_v2c_ncu.log 2>&1; echo "ncu metrics rc=$?"
python tools/lab_metrics.py $O/l2lab_v2c_metrics.csv 2>/dev/null | head -12
timeout 600 ncu --metrics $M --clock-control none -k regex:vcs_light -s 3 -c 1 --csv --log-file $O/light_pairs_metrics.csv python bench_extra.py --only c1 --no-cpu --steps 3 > $O/light_pairs_ncu.log 2>&1; echo "ncu light rc=$?"
python tools/lab_metrics.py $O/light_pairs_metrics.csv 2>/dev/null | head -4
