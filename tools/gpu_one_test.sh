#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "$1" > gpurun_out/one_test.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/one_test.log
