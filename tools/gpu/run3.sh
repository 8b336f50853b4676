#!/bin/bash
# round-2 GPU call 3: whole test suite, tree-walk knobs A/B, REF warp kernel with dynamic distribution, flat bounce occupancy
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02c_tests.log
tail -12 gpurun_out/r02c_tests.log
grep -n "fixed shapes:" gpurun_out/r02c_tests.log
P="python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2"
{
$P $R
$P $R --tune l2_persist=0
$P $R --tune trace_occ=4
$P $R --tune walk_steps=1
$P $R --tune walk_steps=3
$P $R --tune walk_steps=1 --tune trace_occ=4
$P $R --tune leaf_max=16 --tune trace_occ=4
$P $R --tune lanes=2
$P $R --tune lanes=6
$P $R --tune pass_slots=8388608
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
$P --scene CORNELL --spp 64 --frames 3
$P --scene CORNELL --spp 64 --frames 3 --tune bounce_occ=4
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3 --tune bounce_occ=4
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --profile 1 --frames 2
$P --mode REF --scene HEIGHTFIELD --n 708 --frames 2
$P --mode REF --scene CORNELL --frames 2
} > gpurun_out/r02c_timings.log 2>&1
cat gpurun_out/r02c_timings.log
