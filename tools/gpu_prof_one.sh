#!/bin/bash
# ncu --set full capture of one kernel (regex $1, skip $2 launches, out name $3) on the cfg2 bench command
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${4:-}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f $CMD > gpurun_out/ncu_$3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$3.log
