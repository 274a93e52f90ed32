#!/bin/bash
# TMA gather4 probe (tools/gather4_probe.cu, built here with nvcc -gencode arch=compute_100a,code=sm_100a into tools/_bin/)
mkdir -p gpurun_out
timeout 120 tools/_bin/gather4_probe > gpurun_out/gather4_probe.log 2>&1; echo "rc=$?"; cat gpurun_out/gather4_probe.log
