#!/bin/bash
# round-2 GPU call 24: flat diffuse bounce at 4 CTAs / SM with one ray at a time through the pair loops (no spills) vs 3 CTAs dual
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "same_seed or pair_padding or passes_in_flight" > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02x_tests.log
tail -3 gpurun_out/r02x_tests.log | cut -c1-200
P="timeout 120 python tools/profile_run.py"
{
for rep in 1 2; do
$P --scene CORNELL --spp 64 --frames 5
$P --scene CORNELL --spp 64 --frames 5 --tune bounce_occ=4
done
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3
$P --scene CORNELL --w 3840 --h 2160 --spp 64 --depth 8 --frames 3 --tune bounce_occ=4
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3 --tune bounce_occ=4
} > gpurun_out/r02x_timings.log 2>&1
cat gpurun_out/r02x_timings.log | cut -c1-130
