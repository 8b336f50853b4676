#!/bin/bash
# round-2 GPU call 12: first run of the BVH walk (tune walk=3): parity tests, then timings against the octree walk
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "bvh" -s > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02l_tests.log
tail -30 gpurun_out/r02l_tests.log | cut -c1-250
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2"
{
$P $R --tune walk=3 --tune debug_tree=1
$P $R --tune walk=3 --tune walk_steps=4
$P $R --tune walk=3 --tune walk_steps=8
$P $R --tune walk=3 --tune walk_steps=12
$P $R --tune walk=3 --tune trace_occ=3
$P $R --tune walk=3 --tune sort_rays=0
$P $R
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2 --tune walk=3
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
} > gpurun_out/r02l_timings.log 2>&1
cat gpurun_out/r02l_timings.log | cut -c1-200
