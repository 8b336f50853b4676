#!/bin/bash
# final-state check: whole GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r02zx_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zx_tests.log
tail -5 gpurun_out/r02zx_tests.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02zx_bench.json 2> gpurun_out/r02zx_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02zx_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02zx_ref.json 2> gpurun_out/r02zx_ref.err; echo "ref rc=$?"
