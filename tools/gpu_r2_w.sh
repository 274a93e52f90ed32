#!/bin/bash
# Round 2, pass 22: K3M for every query length (ms_max_terms 256) with its own budget for long queries — sweep on the MCP-replay shard
mkdir -p gpurun_out
timeout 900 python tools/ab_opts.py --workload cfg5-shard --batches 4 --out gpurun_out/r02w_ab_cfg5.jsonl --base "overlap=1,ms_max_terms=256,ms_budget_long=70,ms_long_terms=16" \
  --set "" --set "ms_budget_long=40" --set "ms_budget_long=55" --set "ms_budget_long=85" --set "ms_long_terms=8" --set "ms_long_terms=32" --set "ms_budget_long=55,ms_long_terms=8" --set "ms_max_terms=16" \
  2> gpurun_out/r02w_ab_cfg5.err | cut -c1-600; echo "rc=$?"; tail -2 gpurun_out/r02w_ab_cfg5.err
