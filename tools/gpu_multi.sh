#!/bin/bash
# 2-GPU check of the sharded bench path + reference arm
mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"; cat gpurun_out/bench_2gpu.json | cut -c1-1800; tail -5 gpurun_out/bench_2gpu.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
