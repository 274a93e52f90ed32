#!/bin/bash
# Round 2, pass 14: K2T deferred appends (per-warp pending buffer) — GPU suite, then per-segment A/B of the four append
# variants on a resident cfg4 shard (serial per survivor / per-column reservation / always deferred / hybrid = default),
# and K3M stage-ratio sets on the default build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02n_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02n_pytest_gpu.log
for L in default serialpush agg pend16; do
  if [ $L = default ]; then unset VB200_LIB; else export VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_$L.so; fi
  timeout 600 python tools/ab_opts.py --workload cfg4 --batches 10 --out gpurun_out/r02n_ab_cfg4_$L.jsonl --set "overlap=0" \
    2> gpurun_out/r02n_ab_cfg4_$L.err | cut -c1-900; echo "cfg4 $L rc=$?"; tail -2 gpurun_out/r02n_ab_cfg4_$L.err
done
unset VB200_LIB
timeout 900 python tools/ab_opts.py --workload cfg4 --batches 8 --out gpurun_out/r02n_ab_cfg4_stages.jsonl --base "overlap=0,ms_budget=100" \
  --set "" --set "ms_stage_ratio=8" --set "ms_stage_ratio=16" --set "ms_stage_ratio=64" --set "ms_stage_ratio=128" --set "ms_budget=50" --set "ms_budget=70" --set "seg_ratio=4" --set "seg_ratio=16" \
  2> gpurun_out/r02n_ab_cfg4_stages.err | cut -c1-900; echo "stages rc=$?"; tail -2 gpurun_out/r02n_ab_cfg4_stages.err
