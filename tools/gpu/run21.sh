#!/bin/bash
# round-2 GPU call 21: whole suite and bench with the 4-wide BVH walk as the default
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02u_tests.log
tail -8 gpurun_out/r02u_tests.log | cut -c1-250
P="timeout 200 python tools/profile_run.py"
{
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --profile 1 --tune debug_tree=1
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
} > gpurun_out/r02u_timings.log 2>&1
cat gpurun_out/r02u_timings.log | cut -c1-200
python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02u_bench.err
