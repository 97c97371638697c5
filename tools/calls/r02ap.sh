#!/bin/bash
# round 2, call ap (8 GPUs): FINAL build -- the bench line at N = 1, 2, 4, 8 and the reference arm at N = 1, as the driver launches them
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02ap_bench_n1.json 2> gpurun_out/r02ap_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02ap_bench_ref.json 2> gpurun_out/r02ap_bench_ref.err
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r02ap_bench_n$n.json 2> gpurun_out/r02ap_bench_n$n.err
echo "rc=$?" >> gpurun_out/r02ap_bench_n$n.err
done
echo done
