#!/bin/bash
# Round 2, pass 19: K3M work-unit size after the ownership reorder (per-posting work got cheaper: per-unit set-up weighs more)
mkdir -p gpurun_out
for W in cfg4 cfg2; do
timeout 900 python tools/ab_opts.py --workload $W --batches 10 --out gpurun_out/r02t_ab_$W.jsonl --base "overlap=1" \
  --set "" --set "ms_chunk=1024" --set "ms_chunk=2048" --set "ms_chunk=4096" --set "ms_chunk=8192" --set "" \
  2> gpurun_out/r02t_ab_$W.err | cut -c1-700; echo "rc=$?"; tail -3 gpurun_out/r02t_ab_$W.err
done
