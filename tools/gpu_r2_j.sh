#!/bin/bash
# Round 2, pass 10: source-level captures of the B = 1 kernels on cfg1 (K1F, its merge) + launch list
mkdir -p gpurun_out
CMD="python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline --no-api"
$CMD > gpurun_out/r02j_plain.json 2> gpurun_out/r02j_plain.err || { echo plain failed; tail -5 gpurun_out/r02j_plain.err; exit 1; }
cut -c1-400 gpurun_out/r02j_plain.json
ncu --set full --clock-control none --import-source on -k regex:vb_dense_scan1 -s 4 -c 1 -o gpurun_out/r02j_k1f -f $CMD > gpurun_out/ncu_r02j_k1f.log 2>&1; echo "ncu k1f rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vb_compact -s 4 -c 1 -o gpurun_out/r02j_merge -f $CMD > gpurun_out/ncu_r02j_merge.log 2>&1; echo "ncu merge rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02j_launches_cfg1.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep
