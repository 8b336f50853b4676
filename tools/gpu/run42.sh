#!/bin/bash
# camera segment fused for scenes with mirror / glass too: parity + same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py tests/test_analytic.py tests/test_path_link.py -m gpu -q > gpurun_out/r02zw_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zw_tests.log
tail -6 gpurun_out/r02zw_tests.log | cut -c1-220
P="timeout 120 python tools/profile_run.py"
{
for rep in 1 2; do
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 4
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 4 --tune fuse_first=0
done
$P --scene CORNELL --spp 64 --frames 4
} > gpurun_out/r02zw_timings.log 2>&1
cat gpurun_out/r02zw_timings.log | cut -c1-150
