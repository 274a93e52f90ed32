#!/bin/bash
# Round 2, pass 21: K3M term limit on the MCP-replay shard (2..64-term queries) after the ownership reorder
mkdir -p gpurun_out
timeout 900 python tools/ab_opts.py --workload cfg5-shard --batches 4 --out gpurun_out/r02v2_ab_cfg5.jsonl --base "overlap=1,ms_max_terms=16" \
  --set "ms_max_terms=256" --set "ms_max_terms=256,ms_chunk=1024" --set "ms_max_terms=256,ms_stage_ratio=8" --set "ms_max_terms=256,ms_budget=70" 2> gpurun_out/r02v2_ab_cfg5.err | cut -c1-700; echo "rc=$?"; tail -2 gpurun_out/r02v2_ab_cfg5.err
