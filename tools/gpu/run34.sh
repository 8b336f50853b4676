#!/bin/bash
# last vertex folded into the launch that finds it: parity + same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py tests/test_analytic.py tests/test_path_link.py tests/test_fixed_shapes.py -m gpu -q > gpurun_out/r02zj_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zj_tests.log
tail -6 gpurun_out/r02zj_tests.log | cut -c1-220
P="timeout 120 python tools/profile_run.py"
{
for rep in 1 2; do
$P --scene CORNELL --spp 64 --frames 5
$P --scene CORNELL --spp 64 --frames 5 --tune fold_last=0
done
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3 --tune fold_last=0
$P --scene CORNELL --spp 64 --frames 5 --world 8
$P --scene CORNELL --spp 64 --frames 5 --world 8 --tune fold_last=0
} > gpurun_out/r02zj_timings.log 2>&1
cat gpurun_out/r02zj_timings.log | cut -c1-140
