#!/bin/bash
# round-2 GPU call 15: kernel-by-kernel launch list of the room scene with the BVH walk
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --tune walk=3 --frames 2 > gpurun_out/r02o_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02o_launches.csv python tools/profile_run.py --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --tune walk=3 > gpurun_out/r02o_ncu.log 2>&1
cat gpurun_out/r02o_plain.log | cut -c1-200
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --tune walk=3 --profile 1
