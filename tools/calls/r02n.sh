#!/bin/bash
set -x
mkdir -p gpurun_out
QC_PROFILE_SETUP=1 timeout 600 python tools/e2e_api_probe.py 1024 1024 > gpurun_out/r02n_e2e_api.log 2>&1
echo done
