#!/bin/bash
# C2 after the fused camera segment: lanes / pass size sweep (same box)
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
R="--scene CORNELL --spp 64 --frames 5"
{
$P $R
$P $R --tune lanes=2
$P $R --tune lanes=3
$P $R --tune lanes=6
$P $R --tune lanes=8
$P $R --tune pass_slots=4194304
$P $R --tune pass_slots=16777216
$P $R --tune pass_slots=4194304 --tune lanes=8
$P $R --tune pass_slots=16777216 --tune lanes=2
$P $R
} > gpurun_out/r02ze_timings.log 2>&1
cat gpurun_out/r02ze_timings.log | cut -c1-140
