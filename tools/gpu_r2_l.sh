#!/bin/bash
# Round 2, pass 12: aggregated epilogue push (K2 / K2T) — GPU suite, then A/B on one resident cfg4 shard: K3M resident CTAs
# per SM x K2T ring depth (does the sparse chain run UNDER the tensor kernel when it leaves room for it?), segment ratio
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02l_pytest_gpu.log
timeout 900 python tools/ab_opts.py --workload cfg4 --batches 8 --out gpurun_out/r02l_ab_cfg4.jsonl \
  --set "" --set "ms_ctas=6,k2t_stages=3" --set "ms_ctas=4" --set "ms_ctas=7,k2t_stages=3" --set "ms_ctas=3,k2t_stages=3" \
  --set "k2t_stages=3" --set "overlap=0" --set "seg_ratio=16" --set "seg_ratio=32" --set "ms_ctas=6,k2t_stages=3,seg_ratio=16" \
  2> gpurun_out/r02l_ab_cfg4.err | cut -c1-420; echo "ab rc=$?"; tail -3 gpurun_out/r02l_ab_cfg4.err
