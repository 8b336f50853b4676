#!/bin/bash
# pooled builder temporaries: upload timings in the run_configs sequence + tree / path tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tree_build.py tests/test_path_gpu.py tests/test_full_size_gpu.py -m gpu -q > gpurun_out/r02za_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02za_tests.log
tail -4 gpurun_out/r02za_tests.log | cut -c1-200
G19_DEBUG_TREE=1 python tools/run_configs.py --c4-spp 4 --c5-spp 4 > gpurun_out/r02za_configs.md 2> gpurun_out/r02za_upload.log
grep "path_upload\|upload:\|bvh:" gpurun_out/r02za_upload.log | cut -c1-200
grep "C4\|open" gpurun_out/r02za_configs.md | cut -c1-220
