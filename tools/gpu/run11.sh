#!/bin/bash
# round-2 GPU call 11: whole suite, upload timing of the 1 M-triangle scene, full bench line + reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02k_tests.log
tail -6 gpurun_out/r02k_tests.log | cut -c1-250
P="python tools/profile_run.py"
{
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --frames 2 --tune debug_tree=1
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 16 --frames 2 --tune debug_tree=1 --tune upload_threads=1
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2
} > gpurun_out/r02k_timings.log 2>&1
cat gpurun_out/r02k_timings.log | cut -c1-200
python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02k_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_ref.json 2> gpurun_out/r02k_ref.err; echo "ref rc=$?"
