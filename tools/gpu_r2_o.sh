#!/bin/bash
# Round 2, pass 15: compact run-time rare path in the K2 / K2T epilogues (vb_append_flagged) — GPU suite, per-segment A/B
# against the unrolled per-survivor build on resident cfg4 / cfg2 / cfg3-b256 shards
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02o_pytest_gpu.log
for W in cfg4 cfg3-b256-s50 cfg2; do
for L in default serialpush; do
  if [ $L = default ]; then unset VB200_LIB; else export VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_$L.so; fi
  timeout 600 python tools/ab_opts.py --workload $W --batches 10 --out gpurun_out/r02o_ab_${W}_$L.jsonl --set "overlap=0" --set "overlap=1" \
    2> gpurun_out/r02o_ab_${W}_$L.err | cut -c1-900; echo "$W $L rc=$?"; tail -2 gpurun_out/r02o_ab_${W}_$L.err
done; done
