#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/nlevel_retire_ab.py > gpurun_out/r02g_nlevel_ab.log 2>&1
echo done
