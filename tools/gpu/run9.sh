#!/bin/bash
# round-2 GPU call 9: hint-sized ray sort A/B, source-level ncu capture of the flat-scene bounce kernel
mkdir -p gpurun_out
python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py -m gpu -q --maxfail=12 > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02i_tests.log
tail -4 gpurun_out/r02i_tests.log | cut -c1-250
P="python tools/profile_run.py"
{
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 3
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 3 --tune sort_rays=0
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 3
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 3 --tune sort_rays=0
$P --scene CORNELL --spp 64 --frames 3
} > gpurun_out/r02i_timings.log 2>&1
cat gpurun_out/r02i_timings.log | cut -c1-220
$P --scene CORNELL --spp 8 > gpurun_out/r02i_c2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bounce_flat_kernel|raygen_extend_flat' -c 6 -o gpurun_out/r02i_c2 $P --scene CORNELL --spp 8 > gpurun_out/r02i_c2_ncu.log 2>&1
cat gpurun_out/r02i_c2_plain.log | cut -c1-200
