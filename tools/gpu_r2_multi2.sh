#!/bin/bash
# Round 2: N-GPU lines exactly as the driver launches them (torchrun): default workload (cfg4 weak form), reference arm, cfg5
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/r02s_scale_cfg4_g$N.json 2> gpurun_out/r02s_scale_cfg4_g$N.err; echo "cfg4 x$N rc=$?"; cut -c1-250 gpurun_out/r02s_scale_cfg4_g$N.json; grep -o '"e2e": {[^}]*}' gpurun_out/r02s_scale_cfg4_g$N.json | cut -c1-300; tail -3 gpurun_out/r02s_scale_cfg4_g$N.err
timeout 600 $TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/r02s_scale_reference_g$N.json 2> gpurun_out/r02s_scale_reference_g$N.err; echo "reference x$N rc=$?"; cut -c1-200 gpurun_out/r02s_scale_reference_g$N.json
timeout 900 $TR bench.py --gpus $N --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02s_scale_cfg5_g$N.json 2> gpurun_out/r02s_scale_cfg5_g$N.err; echo "cfg5 x$N rc=$?"; cut -c1-250 gpurun_out/r02s_scale_cfg5_g$N.json; grep -o '"e2e": {[^}]*}' gpurun_out/r02s_scale_cfg5_g$N.json | cut -c1-300; tail -3 gpurun_out/r02s_scale_cfg5_g$N.err
