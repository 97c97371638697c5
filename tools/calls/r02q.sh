#!/bin/bash
# round 2, call q (8 GPUs): the bench line at N = 8 and N = 4, as the driver launches it
set -x
mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02q_bench_n$n.json 2> gpurun_out/r02q_bench_n$n.err
echo "rc=$?" >> gpurun_out/r02q_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02q_bench_n8_reference.json 2> gpurun_out/r02q_bench_n8_reference.err
echo done
