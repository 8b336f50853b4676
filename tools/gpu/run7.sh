#!/bin/bash
# round-2 GPU call 7: heavy-ray split in REF mode (tests + A/B), ncu source-level capture of the tree walk at the end state
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_tests.log
tail -8 gpurun_out/r02g_tests.log | cut -c1-250
P="python tools/profile_run.py"
{
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2 --tune ref_heavy=0
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2 --tune ref_heavy=32
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2 --tune ref_heavy=512
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --profile 1 --tune debug_tree=1
$P --mode REF --scene HEIGHTFIELD --n 708 --frames 2
$P --mode REF --scene CORNELL --frames 2
} > gpurun_out/r02g_timings.log 2>&1
cat gpurun_out/r02g_timings.log | cut -c1-220
$P --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 > gpurun_out/r02g_room_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|raygen_extend_kernel' -c 4 -o gpurun_out/r02g_room $P --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 > gpurun_out/r02g_room_ncu.log 2>&1
cat gpurun_out/r02g_room_plain.log | cut -c1-200
