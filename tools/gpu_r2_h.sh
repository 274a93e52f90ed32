#!/bin/bash
# Round 2, pass 8: full GPU suite after the stage-ordering fix; K3H / K3M budget and stage-ratio sweeps.
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f q/s  ms/step %.3f  e2e %.0f  phases/batch %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_batch"].items()}))
except Exception as e:
    print("  no line:", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -8 gpurun_out/r02h_pytest_gpu.log
for BU in 50 30 70 100; do
  VB200_MH_BUDGET=$BU timeout 600 python bench.py --workload cfg5-shard --steps 3 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02h_cfg5_bu$BU.json 2> gpurun_out/r02h_cfg5_bu$BU.err
  echo "cfg5-shard mh_budget=$BU rc=$?"; line gpurun_out/r02h_cfg5_bu$BU.json; tail -2 gpurun_out/r02h_cfg5_bu$BU.err
done
VB200_SPARSE_MH=0 timeout 600 python bench.py --workload cfg5-shard --steps 3 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02h_cfg5_k3.json 2> gpurun_out/r02h_cfg5_k3.err
echo "cfg5-shard K3 for long queries rc=$?"; line gpurun_out/r02h_cfg5_k3.json
for BU in 100 70 50; do
  for SR in 32 8; do
    VB200_MS_BUDGET=$BU VB200_MS_STAGE_RATIO=$SR timeout 600 python bench.py --workload cfg4 --steps 6 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02h_cfg4_bu${BU}_sr$SR.json 2> gpurun_out/r02h_cfg4_bu${BU}_sr$SR.err
    echo "cfg4 ms_budget=$BU stage_ratio=$SR rc=$?"; line gpurun_out/r02h_cfg4_bu${BU}_sr$SR.json; tail -2 gpurun_out/r02h_cfg4_bu${BU}_sr$SR.err
  done
done
for W in cfg1 cfg3-b1-s1 cfg2 cfg3-b256-s50; do
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02h_$W.json 2> gpurun_out/r02h_$W.err
  echo "$W rc=$?"; line gpurun_out/r02h_$W.json; tail -2 gpurun_out/r02h_$W.err
done
