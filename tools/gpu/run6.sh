#!/bin/bash
# round-2 GPU call 6: tests with Philox-7 + REF two-band overlap + 3 interleaved node-test chains; timings; bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -8 gpurun_out/r02f_tests.log | cut -c1-250
P="python tools/profile_run.py"
{
$P --scene CORNELL --spp 64 --frames 3
$P --scene CORNELL --spp 64 --profile 1
$P --scene CORNELL_GLASS --spp 64 --depth 12 --frames 3
$P --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --frames 2
$P --scene HEIGHTFIELD --n 708 --spp 64 --frames 2
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --frames 2
$P --mode REF --scene HEIGHTFIELD_ROOM --n 708 --profile 1 --tune debug_tree=1
$P --mode REF --scene HEIGHTFIELD --n 708 --frames 2
$P --mode REF --scene CORNELL --frames 2
} > gpurun_out/r02f_timings.log 2>&1
cat gpurun_out/r02f_timings.log | cut -c1-220
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02f_bench.err
