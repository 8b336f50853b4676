#!/bin/bash
# round-2 GPU call 8: ray sorting A/B on the tree scenes, PATH tests
mkdir -p gpurun_out
python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py tests/test_path_link.py tests/test_analytic.py -m gpu -q --maxfail=12 > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests.log
tail -8 gpurun_out/r02h_tests.log | cut -c1-250
P="python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2"
{
$P $R
$P $R --tune sort_rays=0
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2 --tune sort_rays=0
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --profile 1
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --profile 1 --tune sort_rays=0
$P $R --tune trace_occ=3
$P $R --tune refill=16
$P $R --tune walk_steps=1
$P $R --tune walk_steps=3
} > gpurun_out/r02h_timings.log 2>&1
cat gpurun_out/r02h_timings.log | cut -c1-220
