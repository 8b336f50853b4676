#!/bin/bash
# round-2 GPU call 2: whole test suite, walk A/B, REF warp kernel timing, bench, ncu of the new tree walk
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b_tests.log
tail -25 gpurun_out/r02b_tests.log
P="python tools/profile_run.py"
{
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=0
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2 --tune walk=1
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2 --tune walk=0
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune leaf_max=16
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune leaf_max=4
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune walk_steps=2
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune walk_steps=8
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune refill=4
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune refill=16
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2 --tune walk=1 --tune coop_leaf=0
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --profile 1 --tune walk=1
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --profile 1 --frames 2
$P --mode REF --scene HEIGHTFIELD --n 708 --profile 1 --frames 2
$P --mode REF --scene CORNELL --profile 1 --frames 2
} > gpurun_out/r02b_timings.log 2>&1
cat gpurun_out/r02b_timings.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02b_bench.err
$P --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 > gpurun_out/r02b_room_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|raygen_extend_kernel' -c 4 -o gpurun_out/r02b_room $P --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 > gpurun_out/r02b_room_ncu.log 2>&1
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 > gpurun_out/r02b_ref_plain.log 2>&1 &&
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'ref_visibility' -c 1 -o gpurun_out/r02b_ref $P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 > gpurun_out/r02b_ref_ncu.log 2>&1
ls -la gpurun_out | grep r02b
