#!/bin/bash
mkdir -p gpurun_out
for W in cfg4-shard cfg5-shard; do
 for P in 10 20 40; do
    VB200_SPARSE_PRUNE_FORCE=1 VB200_SPARSE_PRUNE=$P timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sw2_${W}_$P.json 2> gpurun_out/sw2_${W}_$P.err
    python - $W $P <<'PY'
import json, sys
w, p = sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/sw2_{w}_{p}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(w, "force prune", p, "value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()})
except Exception as e: print("parse fail", w, p, e)
PY
 done
done
