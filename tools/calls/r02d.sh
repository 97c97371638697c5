#!/bin/bash
# round 2, call d (2 GPUs): the 2-rank NCCL batch entry point test, the new np = 1024 form, bench at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02d_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_parity.py -q -m gpu -k "two_rank or tensor_memory_handover" > gpurun_out/r02d_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_edge.py -q -m gpu -k "propagate_host" > gpurun_out/r02d_pytest_host.log 2>&1
python tools/size_sweep.py > gpurun_out/r02d_size_sweep.log 2>&1
python bench.py --gpus 2 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
echo done
