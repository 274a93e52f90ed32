#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_sparse_kernel -s 8 -c 1 -o gpurun_out/prof_sparse $CMD > gpurun_out/ncu3.log 2>&1
echo "sparse full rc=$?"
