#!/bin/bash
# Round 2, pass 16: K2T deferred appends in compact form (vb_park_flagged + out-of-line flush) against the compact direct
# append (-DVB_K2T_DIRECT_APPEND) — GPU suite, per-segment A/B on a resident cfg4 shard, interleaved twice (box drift)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02p_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02p_pytest_gpu.log
for R in 1 2; do
for L in default directappend; do
  if [ $L = default ]; then unset VB200_LIB; else export VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_$L.so; fi
  timeout 600 python tools/ab_opts.py --workload cfg4 --batches 10 --out gpurun_out/r02p_ab_cfg4_${L}_$R.jsonl --set "overlap=0" --set "overlap=1" \
    2> gpurun_out/r02p_ab_cfg4_${L}_$R.err | cut -c1-700; echo "cfg4 $L rc=$?"; tail -2 gpurun_out/r02p_ab_cfg4_${L}_$R.err
done; done
