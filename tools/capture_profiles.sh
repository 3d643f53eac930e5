#!/bin/bash
# tools/capture_profiles.sh — run on a B200 box (gpurun): the bench line, the ncu launch list of the same command, and one
# `--set full` capture per hot kernel, all into gpurun_out/ (post-processed here by tools/collect_profiles.sh).
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/final_bench_c2.json 2> $O/final_bench_c2.err
python bench.py --impl reference --steps 5 --warmup 3 > $O/final_bench_reference.json 2>> $O/final_bench_c2.err
python bench_extra.py --only ntap --no-cpu > $O/final_bench_ntap.json 2>> $O/final_bench_c2.err
python bench_extra.py --only planes --no-cpu > $O/final_bench_planes.json 2>> $O/final_bench_c2.err
# launch list of the same bench command (kernel share of the step)
B200_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/final_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --sustained-seconds 0 > $O/final_launches_c2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
B200_PROFILE=1 $NCU --profile-from-start off -k regex:vcs_lanczos2 -o $O/final_full_lanczos2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --sustained-seconds 0 > $O/final_full_lanczos2.log 2>&1
$NCU -k regex:vcs_light -s 3 -o $O/final_full_light python bench_extra.py --only c1 --no-cpu --steps 3 > $O/final_full_light.log 2>&1
$NCU -k regex:comp_kernel -s 3 -o $O/final_full_comp python bench_extra.py --only c4 --no-cpu --steps 3 > $O/final_full_comp.log 2>&1
$NCU -k regex:vcs_planes_fast -s 10 -o $O/final_full_planes python bench_extra.py --only planes --no-cpu --steps 2 > $O/final_full_planes.log 2>&1
$NCU -k regex:ars_pipe -s 2 -o $O/final_full_ars python bench_extra.py --only c5 --no-cpu --steps 3 > $O/final_full_ars.log 2>&1
# the reports carry the sources (--import-source): keep the headline kernel's, summarise the others here and drop them
# (gpurun merges at most 64 MiB back)
for k in light comp ars planes; do
  python tools/ncu_summary.py $O/final_full_$k.ncu-rep $O/final_${k}_ncu.txt > /dev/null 2>&1 && rm -f $O/final_full_$k.ncu-rep
done
python tools/ncu_summary.py $O/final_full_lanczos2.ncu-rep $O/final_lanczos2_ncu.txt > /dev/null 2>&1
du -sh $O; ls -la $O | tail -24
