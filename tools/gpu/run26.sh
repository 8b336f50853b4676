#!/bin/bash
# where a slow scene upload spends its time: the run_configs sequence with upload timings on stderr
mkdir -p gpurun_out
G19_DEBUG_TREE=1 python tools/run_configs.py --c4-spp 4 --c5-spp 4 > gpurun_out/r02z_configs.md 2> gpurun_out/r02z_upload.log
grep -v "REF visibility launch" gpurun_out/r02z_upload.log | cut -c1-200
grep "C4\|open" gpurun_out/r02z_configs.md | cut -c1-220
