#!/bin/bash
# round-2 GPU call 10: one GPU rendering the tiles of rank 0 of 8 / of 4 (what each rank of an N-GPU run computes): lanes and pass size
mkdir -p gpurun_out
P="python tools/profile_run.py"
{
for W in 8 4 2; do
$P --scene CORNELL --spp 64 --frames 6 --world $W
$P --scene CORNELL --spp 64 --frames 6 --world $W --tune lanes=2
$P --scene CORNELL --spp 64 --frames 6 --world $W --tune lanes=3
$P --scene CORNELL --spp 64 --frames 6 --world $W --tune lanes=1
done
$P --scene CORNELL --spp 64 --frames 6 --world 8 --tune lanes=4 --tune pass_slots=2097152
$P --scene CORNELL --spp 64 --frames 6 --world 8 --tune lanes=8 --tune pass_slots=2097152
$P --scene CORNELL --spp 64 --frames 6 --world 8 --tune lanes=2 --tune pass_slots=16777216
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3 --world 8
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3 --world 1
} > gpurun_out/r02j_timings.log 2>&1
cat gpurun_out/r02j_timings.log | cut -c1-150
