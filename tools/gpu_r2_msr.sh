#!/bin/bash
# Round 2, pass 35: K3M stage growth at cfg2 / cfg3-b1-s1 (fewer posting stages = fewer launches + selects on a latency-bound step)
mkdir -p gpurun_out
for W in cfg2 cfg3-b1-s1; do
timeout 300 python tools/ab_opts.py --workload $W --batches 16 --out gpurun_out/r02msr_ab_$W.jsonl --base "overlap=1,ms_stage_ratio=0" \
  --set "" --set "ms_stage_ratio=128" --set "ms_stage_ratio=512" --set "ms_stage_ratio=4096" --set "" 2> gpurun_out/r02msr_ab_$W.err | cut -c1-120,330-560; echo "$W rc=$?"; tail -1 gpurun_out/r02msr_ab_$W.err
done
