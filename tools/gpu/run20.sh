#!/bin/bash
# round-2 GPU call 20: same-box A/B of the BVH walk variants (each line twice)
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 3 --tune bvh_stack=12"
{
for rep in 1 2; do
$P $R --tune walk=4
$P $R --tune walk=4 --tune bvh_spec=3
$P $R --tune walk=4 --tune bvh_spec=2
$P $R --tune walk=4 --tune bvh_spec=0
$P $R --tune walk=3
done
} > gpurun_out/r02t_timings.log 2>&1
cat gpurun_out/r02t_timings.log | cut -c1-140
