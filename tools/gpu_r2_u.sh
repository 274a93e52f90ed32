#!/bin/bash
# Round 2, pass 20: K3M unit size chosen per stage — GPU suite, then resident-shard timings (auto vs fixed 512) per workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02u_pytest_gpu.log
for W in cfg4 cfg2 cfg3-b256-s50 cfg3-b1-s1 cfg5-shard; do
  timeout 900 python tools/ab_opts.py --workload $W --batches 8 --out gpurun_out/r02u_ab_$W.jsonl --base "overlap=1" \
    --set "" --set "ms_chunk=512" --set "" --set "ms_chunk=512" 2> gpurun_out/r02u_ab_$W.err | cut -c1-600; echo "$W rc=$?"; tail -2 gpurun_out/r02u_ab_$W.err
done
