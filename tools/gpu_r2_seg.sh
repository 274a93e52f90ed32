#!/bin/bash
# Round 2, pass 33: segment schedule with the row selection in (each large segment now carries a count / scatter / copy)
mkdir -p gpurun_out
timeout 600 python tools/ab_opts.py --workload cfg4 --batches 6 --out gpurun_out/r02seg_ab_cfg4.jsonl --base "overlap=1,seg_ratio=0,seg_first=2048" \
  --set "" --set "seg_ratio=16" --set "seg_ratio=32" --set "seg_ratio=64" --set "seg_first=16384,seg_ratio=16" --set "seg_first=8192,seg_ratio=32" --set "" 2> gpurun_out/r02seg_ab_cfg4.err | cut -c1-110,330-640; echo "rc=$?"; tail -2 gpurun_out/r02seg_ab_cfg4.err
