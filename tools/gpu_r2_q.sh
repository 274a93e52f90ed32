#!/bin/bash
# Round 2, pass 17: compact vb_compact_kernel (808 instructions, was 9456), fuse kernel with list rows in shared memory —
# GPU suite, then resident-shard timings of every workload (compare with r02_results.md / r02o / r02p)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02q_pytest_gpu.log
for W in cfg4 cfg2 cfg3-b1-s1 cfg3-b1-s50 cfg1 cfg3-b256-s50; do
  timeout 600 python tools/ab_opts.py --workload $W --batches 20 --out gpurun_out/r02q_ab_$W.jsonl --set "overlap=1" \
    2> gpurun_out/r02q_ab_$W.err | cut -c1-800; echo "$W rc=$?"; tail -2 gpurun_out/r02q_ab_$W.err
done
