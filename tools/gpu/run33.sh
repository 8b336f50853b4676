#!/bin/bash
mkdir -p gpurun_out
G19_DEBUG_TREE=1 python tools/run_configs.py --c4-spp 4 --c5-spp 4 > gpurun_out/r02zh_configs.md 2> gpurun_out/r02zh_upload.log
grep "path_upload\|upload:\|bvh:\|REF view" gpurun_out/r02zh_upload.log | grep -v " 8 prim\|0.0 ms, tree\| 14 ent\| 3 ent" | cut -c1-200
grep "C4\|open" gpurun_out/r02zh_configs.md | cut -c1-220
