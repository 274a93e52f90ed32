#!/bin/bash
# Round 2, pass 29: large GEMM segments after the K3M stages (dense_wait_sparse) vs co-running
mkdir -p gpurun_out
for W in cfg4 cfg5-shard cfg3-b256-s50; do
timeout 600 python tools/ab_opts.py --workload $W --batches 6 --out gpurun_out/r02ord_ab_$W.jsonl --base "overlap=1" \
  --set "dense_wait_sparse=0" --set "dense_wait_sparse=1" --set "dense_wait_sparse=0" --set "dense_wait_sparse=1" 2> gpurun_out/r02ord_ab_$W.err | cut -c1-100,330-800; echo "$W rc=$?"; tail -2 gpurun_out/r02ord_ab_$W.err
done
