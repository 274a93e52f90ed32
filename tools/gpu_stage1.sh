#!/bin/bash
# First GPU bring-up: K1-only parity, then the tcgen05 path on its own under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?"
VB200_DENSE_PATH=1 timeout 900 python -m pytest tests -m gpu -x -q \
   -k "not batch_paths and not per_query_filters and not dimensions and not golden" > gpurun_out/t_k1.log 2>&1
echo "k1 tests rc=$?"; tail -5 gpurun_out/t_k1.log
VB200_DENSE_PATH=1 timeout 300 python -m pytest tests/test_gpu_golden.py -x -q -k "0" > gpurun_out/t_golden_k1.log 2>&1
echo "golden k1 rc=$?"; tail -5 gpurun_out/t_golden_k1.log
timeout 180 python -m pytest tests/test_gpu_engine.py -x -q -k "batch_paths" > gpurun_out/t_k2.log 2>&1
echo "k2 tests rc=$?"; tail -15 gpurun_out/t_k2.log
timeout 180 python -m pytest tests/test_gpu_engine.py -x -q -k "dimensions" > gpurun_out/t_dims.log 2>&1
echo "dims rc=$?"; tail -8 gpurun_out/t_dims.log
