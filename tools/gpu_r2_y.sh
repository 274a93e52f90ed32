#!/bin/bash
# Round 2, pass 24: GPU suite with the device-query hand-off and K3M for every query length; bench lines at HEAD
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02y_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/r02y_pytest_gpu.log
for W in cfg4 cfg5-shard cfg3-b1-s1 cfg1; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02y_bench_$W.json 2> gpurun_out/r02y_bench_$W.err; echo "$W rc=$?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02y_bench_$W.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d.get("phases_ms") or d.get("roofline",{}).get("frac"))
except Exception as e: print("parse", e)
P
  tail -2 gpurun_out/r02y_bench_$W.err
done
