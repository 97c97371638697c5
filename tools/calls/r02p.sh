#!/bin/bash
# round 2, call p (2 GPUs): 2-rank NCCL tests and the bench line at N = 2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -m gpu > gpurun_out/r02p_pytest_multirank.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02p_bench_n2.json 2> gpurun_out/r02p_bench_n2.err
echo "rc=$?" >> gpurun_out/r02p_bench_n2.err
echo done
