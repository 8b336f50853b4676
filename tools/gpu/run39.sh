#!/bin/bash
# one rank's eighth of C2 at the end state: lanes / pass sizes (same box)
mkdir -p gpurun_out
P="timeout 120 python tools/profile_run.py"
R="--scene CORNELL --spp 64 --frames 8 --world 8"
{
$P $R
$P $R --tune lanes=2
$P $R --tune lanes=3
$P $R --tune lanes=1
$P $R --tune lanes=8 --tune pass_slots=2097152
$P $R --tune lanes=6 --tune pass_slots=2800000
$P $R --tune lanes=4 --tune pass_slots=2097152
$P $R
$P --scene CORNELL --spp 64 --frames 8
} > gpurun_out/r02zr_timings.log 2>&1
cat gpurun_out/r02zr_timings.log | cut -c1-150
