#!/bin/bash
# tools/capture_profiles.sh — run on a B200 box (gpurun): benchmark lines, ncu launch list of the
# headline bench and one `--set full` capture per hot kernel, all into gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
python bench_extra.py > $O/bench_extra.json 2> $O/bench_extra.err
python bench_extra.py --only c4 --background 3 --no-cpu >> $O/bench_extra.json 2>> $O/bench_extra.err
# launch list of the same bench command (kernel share of the step)
B200_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/launches_c2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
B200_PROFILE=1 $NCU --profile-from-start off -k regex:vcs_lanczos2 -o $O/full_lanczos2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/full_lanczos2.log 2>&1
$NCU -k regex:vcs_light -s 3 -o $O/full_light python bench_extra.py --only c1 --no-cpu --steps 3 > $O/full_light.log 2>&1
$NCU -k regex:comp_kernel -s 3 -o $O/full_comp python bench_extra.py --only c4 --no-cpu --steps 3 > $O/full_comp.log 2>&1
$NCU -k regex:ars_tile -s 2 -o $O/full_ars python bench_extra.py --only c5 --no-cpu --steps 3 > $O/full_ars.log 2>&1
ls -la $O
