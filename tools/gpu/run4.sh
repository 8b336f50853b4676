#!/bin/bash
# round-2 GPU call 4: whole test suite after the stream-hint fix and the shared specular array, knob A/B, memory footprint
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02d_tests.log
tail -12 gpurun_out/r02d_tests.log | cut -c1-250
grep -n "fixed shapes:" gpurun_out/r02d_tests.log | cut -c1-1200
P="python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2"
{
$P $R
$P $R --tune l2_persist=0
$P $R --tune trace_occ=4
$P $R --tune trace_occ=4 --tune l2_persist=0
$P $R --tune walk_steps=1
$P $R --tune walk_steps=3
$P $R --tune lanes=2
$P $R --tune debug_tree=1
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
$P --scene CORNELL --spp 64 --frames 3
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3
$P --scene CORNELL --w 3840 --h 2160 --spp 256 --depth 8 --frames 2
nvidia-smi --query-gpu=memory.used --format=csv
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2
$P --mode REF --scene HEIGHTFIELD --n 708 --frames 2
} > gpurun_out/r02d_timings.log 2>&1
cat gpurun_out/r02d_timings.log | cut -c1-230
