#!/bin/bash
# Round 2, pass 23: K3M budgets after the ownership reorder: short queries (cfg4 / cfg2 / cfg3-b256), long queries (cfg5 shard)
mkdir -p gpurun_out
for W in cfg4 cfg2 cfg3-b256-s50; do
timeout 900 python tools/ab_opts.py --workload $W --batches 8 --out gpurun_out/r02x_ab_$W.jsonl --base "overlap=1,ms_budget=100" \
  --set "" --set "ms_budget=90" --set "ms_budget=80" --set "ms_budget=70" --set "" 2> gpurun_out/r02x_ab_$W.err | cut -c1-100,330-700; echo "$W rc=$?"; tail -2 gpurun_out/r02x_ab_$W.err
done
timeout 900 python tools/ab_opts.py --workload cfg5-shard --batches 4 --out gpurun_out/r02x_ab_cfg5.jsonl --base "overlap=1,ms_max_terms=256,ms_budget_long=85,ms_long_terms=16,ms_budget=100" \
  --set "" --set "ms_budget_long=92" --set "ms_budget_long=80" --set "ms_budget_long=85,ms_budget=85" --set "ms_budget_long=85,ms_stage_ratio=8" --set "ms_budget_long=85,ms_stage_ratio=128" \
  2> gpurun_out/r02x_ab_cfg5.err | cut -c1-100,330-700; echo "rc=$?"; tail -2 gpurun_out/r02x_ab_cfg5.err
