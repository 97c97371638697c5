#!/bin/bash
# round 2, call au (2 GPUs): final bench lines at N = 1 and N = 2 with the gated pipeline, 2-rank NCCL test
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02au_bench_n1.json 2> gpurun_out/r02au_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02au_bench_n2.json 2> gpurun_out/r02au_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_bench_contract.py -q -m gpu > gpurun_out/r02au_pytest.log 2>&1
echo done
