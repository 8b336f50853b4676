#!/bin/bash
# BVH walk: leave the node visits early when many lanes wait for the leaf step (bvh_spec = 1 + 4 * lanes)
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 3"
{
$P $R
$P $R --tune bvh_spec=17
$P $R --tune bvh_spec=33
$P $R --tune bvh_spec=49
$P $R --tune bvh_spec=65
$P $R --tune bvh_spec=33 --tune walk_steps=8
$P $R --tune bvh_spec=49 --tune walk_steps=8
$P $R --tune bvh_spec=49 --tune walk_steps=12
$P $R
} > gpurun_out/r02zb_timings.log 2>&1
cat gpurun_out/r02zb_timings.log | cut -c1-150
