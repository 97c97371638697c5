#!/bin/bash
# round 2, call i: level-skew probe of kernel A; DMMA pipe probe
set -x
mkdir -p gpurun_out
timeout 120 tools/dmma_probe > gpurun_out/r02i_dmma_probe.log 2>&1
timeout 300 python tools/skew_probe.py > gpurun_out/r02i_skew_probe.log 2>&1
echo done
