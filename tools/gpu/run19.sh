#!/bin/bash
# round-2 GPU call 19: shared-memory stack depth of the BVH walks (the L1 is what the carve-out leaves)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "bvh" > gpurun_out/r02s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02s_tests.log
tail -4 gpurun_out/r02s_tests.log | cut -c1-250
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2"
{
$P $R --tune walk=4
$P $R --tune walk=4 --tune bvh_stack=8
$P $R --tune walk=4 --tune bvh_stack=12
$P $R --tune walk=4 --tune bvh_stack=16
$P $R --tune walk=4 --tune bvh_stack=32
$P $R --tune walk=4 --tune bvh_stack=16 --tune walk_steps=4
$P $R --tune walk=4 --tune bvh_stack=16 --tune walk_steps=6
$P $R --tune walk=3 --tune bvh_stack=16
$P $R --tune walk=3
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2 --tune walk=4 --tune bvh_stack=16
} > gpurun_out/r02s_timings.log 2>&1
cat gpurun_out/r02s_timings.log | cut -c1-200
