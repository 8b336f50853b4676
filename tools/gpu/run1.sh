#!/bin/bash
# round-2 GPU call 1: tests, default bench line, baseline ncu captures of the tree kernels and of REF visibility
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02a_tests.log
tail -5 gpurun_out/r02a_tests.log
python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
python tools/profile_run.py --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 --profile 1 > gpurun_out/r02a_room_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_kernel|raygen_extend_kernel' -c 6 -o gpurun_out/r02a_room python tools/profile_run.py --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --spp 32 > gpurun_out/r02a_room_ncu.log 2>&1
cat gpurun_out/r02a_room_plain.log
python tools/profile_run.py --mode REF --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 --profile 1 > gpurun_out/r02a_ref_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ref_visibility' -c 1 -o gpurun_out/r02a_ref python tools/profile_run.py --mode REF --scene HEIGHTFIELD_ROOM --n 708 --w 960 --h 540 > gpurun_out/r02a_ref_ncu.log 2>&1
cat gpurun_out/r02a_ref_plain.log
python tools/profile_run.py --mode REF --scene HEIGHTFIELD_ROOM --n 708 --profile 1
python tools/profile_run.py --scene HEIGHTFIELD_ROOM --n 708 --spp 64 --profile 0 --frames 2
python tools/profile_run.py --scene HEIGHTFIELD --n 708 --spp 64 --profile 0 --frames 2
ls -la gpurun_out | tail -12
