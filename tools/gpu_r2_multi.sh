#!/bin/bash
# Round 2: N-GPU lines exactly as the driver launches them (torchrun, one rank per GPU): the default workload (cfg4 weak
# form), cfg5 (MCP replay) and the reference arm under torchrun (explicit OpenMP thread count).  Usage: gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02m_pytest_gpu.log
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/r02_scale_cfg4_g$N.json 2> gpurun_out/r02_scale_cfg4_g$N.err; echo "cfg4 x$N rc=$?"; cut -c1-700 gpurun_out/r02_scale_cfg4_g$N.json; tail -3 gpurun_out/r02_scale_cfg4_g$N.err
timeout 900 $TR bench.py --gpus $N --impl reference --steps 5 --warmup 3 > gpurun_out/r02_scale_reference_g$N.json 2> gpurun_out/r02_scale_reference_g$N.err; echo "reference x$N rc=$?"; cut -c1-300 gpurun_out/r02_scale_reference_g$N.json
timeout 900 $TR bench.py --gpus $N --workload cfg5 --steps 3 --warmup 3 > gpurun_out/r02_scale_cfg5_g$N.json 2> gpurun_out/r02_scale_cfg5_g$N.err; echo "cfg5 x$N rc=$?"; cut -c1-700 gpurun_out/r02_scale_cfg5_g$N.json; tail -3 gpurun_out/r02_scale_cfg5_g$N.err
timeout 900 $TR bench.py --gpus $N --workload cfg2 --steps 10 --warmup 5 > gpurun_out/r02_scale_cfg2_g$N.json 2> gpurun_out/r02_scale_cfg2_g$N.err; echo "cfg2 x$N rc=$?"; cut -c1-500 gpurun_out/r02_scale_cfg2_g$N.json; tail -3 gpurun_out/r02_scale_cfg2_g$N.err
