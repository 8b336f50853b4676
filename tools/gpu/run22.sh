#!/bin/bash
# round-2 multi-GPU call: N-rank bit-identity (tests/test_multi_gpu.py, tools/dist_check.py) and bench.py under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r02zu_n${N}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02zu_n${N}_tests.log
tail -5 gpurun_out/r02zu_n${N}_tests.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02zu_n${N}_bench.json 2> gpurun_out/r02zu_n${N}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02zu_n${N}_bench.err
tail -c 300 gpurun_out/r02zu_n${N}_bench.json
