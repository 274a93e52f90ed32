#!/bin/bash
# Round 2, first GPU pass: GPU test suite, then K3M (posting-driven MaxScore) vs K3 A/B on the batched workloads.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/r02a_pytest_gpu.log
for W in cfg2 cfg3-b256-s50 cfg4-shard cfg5-shard cfg3-b1-s50; do
  ST=200; [ "$W" = "cfg3-b256-s50" ] && ST=20; [ "$W" = "cfg4-shard" ] && ST=6; [ "$W" = "cfg5-shard" ] && ST=3
  for MS in 1 0; do
    [ "$MS" = "0" ] && [ "$W" != "cfg2" ] && [ "$W" != "cfg3-b256-s50" ] && continue
    VB200_SPARSE_MS=$MS timeout 900 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_${W}_ms$MS.json 2> gpurun_out/r02a_bench_${W}_ms$MS.err
    echo "$W ms=$MS rc=$?"
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02a_bench_${W}_ms$MS.json").read().strip().splitlines()[-1])
    print("  value %.0f q/s  ms/step %.3f  e2e %.0f  phases %s  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_step"].items()}, d.get("parity_spot_check")))
except Exception as e:
    print("  no line:", e)
PY
  done
done
