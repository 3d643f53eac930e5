#!/bin/bash
# First device run of the next round: everything that was written after the round-1 GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_first.sh'
# Each step has its own timeout and writes under gpurun_out/; a red step does not stop the later ones.
set -u
mkdir -p gpurun_out
export B200_TEST_EXPERIMENTAL=1
run() { local name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/nr_$name.log" 2>&1; echo "rc=$? ($name)"; tail -4 "gpurun_out/nr_$name.log"; }
# 1. packed RGB -> 4:2:0, RGB -> RGB, 4:2:2 / 4:4:4 inputs (opt-in product paths: B200_VCS_EXPERIMENTAL is set by the tests)
run rgbin 240 python -m pytest tests/test_vcs_rgbin_gpu.py -q -p no:cacheprovider
# 2. destination rectangle + borders (new vcs_border_kernel around the existing kernels)
run borders 180 python -m pytest tests/test_vcs_borders_gpu.py -q -p no:cacheprovider
# 2b. audioresample's method / filter-mode / interpolation properties (mostly host tables in front of the default kernels; new
#     device code: the two-row linear blend of the interpolated mode and ars_small_kernel for nearest / linear / cubic)
run arsopts 240 python -m pytest tests/test_ars_options_gpu.py -q -p no:cacheprovider
# 2c. tensor-path variant of the 2:1 kernel: parity, then the headline bench with it (compare with the default line)
run l2mma 240 python -m pytest tests/test_vcs_l2mma_gpu.py -q -p no:cacheprovider
B200_L2_X4=1 run l2x4 300 python -m pytest tests/test_vcs_lanczos2_gpu.py -q -p no:cacheprovider
B200_L2_MMA=1 timeout 300 python bench.py > gpurun_out/nr_bench_l2mma.json 2> gpurun_out/nr_bench_l2mma.err; tail -c 600 gpurun_out/nr_bench_l2mma.json
B200_L2_X4=1 timeout 300 python bench.py > gpurun_out/nr_bench_x4.json 2> gpurun_out/nr_bench_x4.err; tail -c 600 gpurun_out/nr_bench_x4.json
timeout 300 python bench.py > gpurun_out/nr_bench_default.json 2> gpurun_out/nr_bench_default.err; tail -c 600 gpurun_out/nr_bench_default.json
# 3. random sweep of the RGB -> 4:2:0 path, every outcome recorded in gpurun_out/cross_check.json
run rgbsweep 90 python tools/gpu_cross_check.py 60 rgb
# 4. the regular device suite (the generic kernel and the plan builder changed underneath it)
unset B200_TEST_EXPERIMENTAL
run suite 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider
