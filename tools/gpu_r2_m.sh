#!/bin/bash
# Round 2, pass 13: K3M ownership-after-score, K2/K2T aggregated push — GPU suite, then per-segment A/B on resident shards:
# default library vs the -DVB_PUSH_SERIAL build (voitta-rag_b200/build.py build_variant), cfg4 / cfg2 / cfg3-b256-s50
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02m_pytest_gpu.log
for W in cfg4 cfg2 cfg3-b256-s50; do
  for L in default serialpush; do
    if [ $L = default ]; then unset VB200_LIB; else export VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_$L.so; fi
    timeout 600 python tools/ab_opts.py --workload $W --batches 10 --out gpurun_out/r02m_ab_${W}_$L.jsonl --set "" --set "overlap=0" \
      2> gpurun_out/r02m_ab_${W}_$L.err | cut -c1-900; echo "$W $L rc=$?"; tail -2 gpurun_out/r02m_ab_${W}_$L.err
  done
done
