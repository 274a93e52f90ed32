#!/bin/bash
# iteration check: GPU tests, then bench lines for the listed workloads (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/it_pytest.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/it_pytest.log
for W in ${@:-cfg2}; do
  ST=300; [ "$W" = "cfg3-b256-s50" ] && ST=20; [ "$W" = "cfg4-shard" ] && ST=5; [ "$W" = "cfg5-shard" ] && ST=3; [ "$W" = "cfg3-b1-s50" ] && ST=100
  timeout 900 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/it_$W.json 2> gpurun_out/it_$W.err; echo "$W rc=$?"
  python - $W <<'PY'
import json, sys
w = sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/it_{w}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(w, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()}, {k: (round(v, 4) if v else v) for k, v in r["big_launch"].items()}, "tflops", r["dense_tflops_per_step"])
except Exception as e: print("parse fail", e); print(open(f"gpurun_out/it_{w}.err").read()[-2000:])
PY
done
