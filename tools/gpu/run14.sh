#!/bin/bash
# round-2 GPU call 14: BVH walk with compacted nodes, shared-memory stack, speculative traversal
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "bvh" > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02n_tests.log
tail -5 gpurun_out/r02n_tests.log | cut -c1-250
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=3"
{
$P $R --tune debug_tree=1
$P $R --tune bvh_spec=0
$P $R --tune walk_steps=4
$P $R --tune walk_steps=6
$P $R --tune walk_steps=12
$P $R --tune walk_steps=16
$P $R --tune bvh_leaf=2
$P $R --tune bvh_leaf=8
$P $R --tune refill=16
$P $R --tune sort_rays=0
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2 --tune walk=3
} > gpurun_out/r02n_timings.log 2>&1
cat gpurun_out/r02n_timings.log | cut -c1-200
