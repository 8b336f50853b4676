#!/bin/bash
# final state: lanes sweep with frames in a row (one GPU and one rank's eighth)
mkdir -p gpurun_out
A="timeout 120 python tools/async_frames.py"
{
for L in 4 5 6 8; do
$A --frames 20 --tune lanes=$L
$A --world 8 --frames 100 --tune lanes=$L
done
$A --world 8 --frames 100 --tune lanes=8 --tune pass_slots=2097152
$A --scene CORNELL_GLASS --depth 12 --frames 10 --tune lanes=6
$A --scene CORNELL_GLASS --depth 12 --frames 10
$A --scene HEIGHTFIELD_ROOM --n 708 --frames 3
$A --scene HEIGHTFIELD_ROOM --n 708 --frames 3 --tune overlap_frames=0
} > gpurun_out/r02zv_timings.log 2>&1
cat gpurun_out/r02zv_timings.log | cut -c1-150
