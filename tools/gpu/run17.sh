#!/bin/bash
# round-2 GPU call 17: ncu capture of the quantised 4-wide BVH walk
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
$P --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 --tune walk=4 > gpurun_out/r02q_room_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|raygen_extend_kernel|bounce_kernel' -c 6 -o gpurun_out/r02q_room python tools/profile_run.py --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 --tune walk=4 > gpurun_out/r02q_room_ncu.log 2>&1
cat gpurun_out/r02q_room_plain.log | cut -c1-200
