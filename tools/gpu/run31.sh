#!/bin/bash
# fused camera segment + first bounce (diffuse-only flat scenes): parity tests and same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py tests/test_analytic.py tests/test_path_link.py tests/test_fixed_shapes.py -m gpu -q > gpurun_out/r02zd_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zd_tests.log
tail -6 gpurun_out/r02zd_tests.log | cut -c1-220
P="timeout 120 python tools/profile_run.py"
{
for rep in 1 2; do
$P --scene CORNELL --spp 64 --frames 5
$P --scene CORNELL --spp 64 --frames 5 --tune fuse_first=0
done
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3 --tune fuse_first=0
$P --scene CORNELL --spp 64 --frames 5 --world 8
$P --scene CORNELL --spp 64 --frames 5 --world 8 --tune fuse_first=0
} > gpurun_out/r02zd_timings.log 2>&1
cat gpurun_out/r02zd_timings.log | cut -c1-140
