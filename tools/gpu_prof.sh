#!/bin/bash
# ncu evidence for the cfg2 bench line: per-launch durations and full sections of the top kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/launches_cfg2.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_dense_gemm -s 6 -c 3 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu2.log 2>&1
echo "gemm full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_sparse_kernel -s 6 -c 3 -o gpurun_out/prof_sparse $CMD > gpurun_out/ncu3.log 2>&1
echo "sparse full rc=$?"
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
