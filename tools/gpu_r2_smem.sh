#!/bin/bash
# Round 2, pass 34: room for the other chain's small kernels beside the persistent GEMM CTA (shared memory):
# resident K2 at cfg2 (VB200_K2_SMEM_KB), K2T at cfg4 (k2t_stages 3 instead of 4 frees 48 KB)
mkdir -p gpurun_out
for KB in 0 195 170; do
  VB200_K2_SMEM_KB=$KB timeout 300 python tools/ab_opts.py --workload cfg2 --batches 16 --out gpurun_out/r02smem_ab_cfg2_$KB.jsonl --base "overlap=1" --set "" --set "" 2> gpurun_out/r02smem_ab_cfg2_$KB.err | cut -c1-100; echo "cfg2 smem $KB rc=$?"
done
timeout 300 python tools/ab_opts.py --workload cfg4 --batches 6 --out gpurun_out/r02smem_ab_cfg4.jsonl --base "overlap=1,k2t_stages=0" --set "" --set "k2t_stages=3" --set "" --set "k2t_stages=3" 2> gpurun_out/r02smem_ab_cfg4.err | cut -c1-110; echo "cfg4 rc=$?"
