#!/bin/bash
# round-2 GPU call 5: tests, top-table A/B, REF per-band timing and heaviest ray
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02e_tests.log
tail -6 gpurun_out/r02e_tests.log | cut -c1-250
P="python tools/profile_run.py"
R="--scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2"
{
$P $R
$P $R --tune top_level=0
$P $R --tune top_level=5
$P $R --tune top_level=7
$P $R --tune top_level=7 --tune refill=4
$P $R --tune top_level=7 --tune refill=16
$P $R --tune top_level=7 --tune leaf_max=16
$P $R --tune top_level=7 --tune lanes=2
$P $R --tune top_level=7 --tune walk_steps=1
$P $R --tune top_level=7 --tune walk_steps=3
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --profile 1
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --profile 1 --tune debug_tree=1
$P --mode REF --scene HEIGHTFIELD --n 708 --profile 1 --tune debug_tree=1
} > gpurun_out/r02e_timings.log 2>&1
cat gpurun_out/r02e_timings.log | cut -c1-200
