#!/bin/bash
# frames in a row: overlap on the side lanes -- bit-identity test and async timings (one GPU / one rank's eighth)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py tests/test_full_size_gpu.py -m gpu -q > gpurun_out/r02zs_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zs_tests.log
tail -6 gpurun_out/r02zs_tests.log | cut -c1-220
A="timeout 120 python tools/async_frames.py"
{
for rep in 1 2; do
$A --world 8 --frames 100
$A --world 8 --frames 100 --tune overlap_frames=0
done
$A --world 4 --frames 60
$A --world 4 --frames 60 --tune overlap_frames=0
$A --frames 20
$A --frames 20 --tune overlap_frames=0
$A --scene CORNELL_GLASS --depth 12 --frames 10
$A --scene CORNELL_GLASS --depth 12 --frames 10 --tune overlap_frames=0
} > gpurun_out/r02zs_timings.log 2>&1
cat gpurun_out/r02zs_timings.log | cut -c1-150
