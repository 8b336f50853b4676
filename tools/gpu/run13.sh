#!/bin/bash
# round-2 GPU call 13: BVH walk knobs and an ncu capture (with source) of its trace_kernel
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=3 --tune walk_steps=8"
{
$P $R
$P $R --tune bvh_leaf=2
$P $R --tune bvh_leaf=8
$P $R --tune bvh_leaf=1
$P $R --tune refill=4
$P $R --tune refill=16
$P $R --tune refill=24
$P $R --tune lanes=2
$P $R --tune pass_slots=8388608
} > gpurun_out/r02m_timings.log 2>&1
cat gpurun_out/r02m_timings.log | cut -c1-200
$P --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 --tune walk=3 --tune walk_steps=8 > gpurun_out/r02m_room_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|raygen_extend_kernel' -c 4 -o gpurun_out/r02m_room python tools/profile_run.py --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 --tune walk=3 --tune walk_steps=8 > gpurun_out/r02m_room_ncu.log 2>&1
cat gpurun_out/r02m_room_plain.log | cut -c1-200
