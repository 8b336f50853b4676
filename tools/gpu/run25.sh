#!/bin/bash
# round-2 end-state evidence: smoke, launch list of one bench step, full capture of one pass at the bench's pass size
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/r02zl_plain.json 2> gpurun_out/r02zl_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02zl_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-others > gpurun_out/r02zl_nculist.log 2>&1
tail -c 200 gpurun_out/r02zl_plain.json
python tools/profile_run.py --spp 4 --spp-per-pass 4 > gpurun_out/r02zl_prof_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -o gpurun_out/r02zl_pass python tools/profile_run.py --spp 4 --spp-per-pass 4 > gpurun_out/r02zl_ncufull.log 2>&1
cat gpurun_out/r02zl_prof_plain.log | cut -c1-160
python tools/run_configs.py > gpurun_out/r02zl_configs.md 2>&1; tail -12 gpurun_out/r02zl_configs.md | cut -c1-200
