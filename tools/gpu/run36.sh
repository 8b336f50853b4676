#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ref_fullsize_gpu.py tests/test_full_size_gpu.py tests/test_tree_build.py tests/test_ref_gpu.py -m gpu -q > gpurun_out/r02zo_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zo_tests.log
tail -4 gpurun_out/r02zo_tests.log | cut -c1-200
G19_DEBUG_TREE=1 python tools/run_configs.py --c4-spp 4 --c5-spp 4 > gpurun_out/r02zo_configs.md 2> gpurun_out/r02zo_upload.log
grep "upload:" gpurun_out/r02zo_upload.log | grep "side by side" | cut -c1-200
grep "C4\|open" gpurun_out/r02zo_configs.md | cut -c1-220
