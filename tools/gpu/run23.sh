#!/bin/bash
# round-2 GPU call 23: big primitives through the pair loops, deep-stack test, PATH suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py tests/test_tree_build.py tests/test_path_link.py tests/test_analytic.py tests/test_fixed_shapes.py -m gpu -q > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests.log
tail -12 gpurun_out/r02w_tests.log | cut -c1-250
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 3"
{
$P $R
$P $R
$P $R --tune walk=3
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 3
} > gpurun_out/r02w_timings.log 2>&1
cat gpurun_out/r02w_timings.log | cut -c1-150
