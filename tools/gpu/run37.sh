#!/bin/bash
mkdir -p gpurun_out
for c in c3 c4 c5; do
python bench.py --config $c --steps 2 --warmup 1 --no-cpu --no-others > gpurun_out/r02zp_$c.json 2> gpurun_out/r02zp_$c.err; echo "$c rc=$?"
tail -c 200 gpurun_out/r02zp_$c.err
done
