#!/bin/bash
# tools/audio_ab.sh — device session for the audio pipeline: parity tests, A/B of the two pipeline forms on C5, ncu counters.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_ars_gpu.py tests/test_ars_options_gpu.py tests/test_host_paths_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
echo "== new"; timeout 300 python bench_extra.py --only c5 --no-cpu --steps 10 2>&1 | tail -1 | cut -c1-600 | tee $O/c5_new_20s.json
echo "== rows"; B200_ARS_NO_TENSORMAP=1 timeout 300 python bench_extra.py --only c5 --no-cpu --steps 10 2>&1 | tail -1 | cut -c1-600 | tee $O/c5_rows_20s.json
echo "== v1";  B200_ARS_PIPE_V1=1 timeout 300 python bench_extra.py --only c5 --no-cpu --steps 10 2>&1 | tail -1 | cut -c1-600 | tee $O/c5_v1_20s.json
echo "== new 600 s"; timeout 600 python bench_extra.py --only c5 --no-cpu --steps 3 --seconds 600 2>&1 | tail -1 | cut -c1-600 | tee $O/c5_new_600s.json
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:ars_pipe -s 2 -c 1 --csv --log-file $O/ars_pipe2_metrics.csv python bench_extra.py --only c5 --no-cpu --steps 3 > $O/ars_pipe2_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $O/ars_pipe2_metrics.csv | cut -c1-300
