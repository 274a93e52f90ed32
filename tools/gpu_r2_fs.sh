#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "full_size_cfg4" > gpurun_out/r02fs_pytest.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02fs_pytest.log
