#!/bin/bash
# Round 2, pass 18: segment / stage ratios on a resident cfg4 shard with the new epilogue (appends are cheap now)
mkdir -p gpurun_out
timeout 900 python tools/ab_opts.py --workload cfg4 --batches 10 --out gpurun_out/r02r_ab_cfg4.jsonl --base "ms_budget=100,overlap=1" \
  --set "" --set "seg_ratio=16" --set "seg_ratio=32" --set "seg_ratio=4" --set "ms_stage_ratio=128" --set "ms_stage_ratio=1024" --set "seg_ratio=16,ms_stage_ratio=128" --set "seg_first=4096" --set "seg_first=8192,seg_ratio=16" --set "" \
  2> gpurun_out/r02r_ab_cfg4.err | cut -c1-700; echo "rc=$?"; tail -3 gpurun_out/r02r_ab_cfg4.err
