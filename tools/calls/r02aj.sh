#!/bin/bash
# round 2, call aj (8 GPUs): final build -- 2-rank NCCL tests, the bench line at N = 8 and N = 2 as the driver launches it
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -m gpu > gpurun_out/r02aj_pytest_multirank.log 2>&1
for n in 8 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02aj_bench_n$n.json 2> gpurun_out/r02aj_bench_n$n.err
echo "rc=$?" >> gpurun_out/r02aj_bench_n$n.err
done
echo done
