#!/bin/bash
# sweep of the MaxScore budget (VB200_SPARSE_PRUNE, % of tau) on cfg2 and cfg3-b256-s50
mkdir -p gpurun_out
for P in 0 5 10 20 35 50; do
  for W in cfg2 cfg3-b256-s50; do
    ST=200; [ "$W" = "cfg3-b256-s50" ] && ST=10
    VB200_SPARSE_PRUNE=$P timeout 600 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/sw_${W}_$P.json 2> gpurun_out/sw_${W}_$P.err
    python - $W $P <<'PY'
import json, sys
w, p = sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/sw_{w}_{p}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(w, "prune", p, "value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()})
except Exception as e: print("parse fail", w, p, e)
PY
  done
done
