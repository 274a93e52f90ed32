#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "tiled or batch" 2>&1 | tail -2
for W in cfg2-g8shard cfg3-b256-s50; do
 for SF in 2048 8192; do for SR in 16 32 64; do
    ST=100; [ "$W" = "cfg3-b256-s50" ] && ST=10
    VB200_SEG_FIRST=$SF VB200_SEG_RATIO=$SR timeout 600 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/sw3.json 2> gpurun_out/sw3.err
    python - $W $SF $SR <<'PY'
import json, sys
w, a, b = sys.argv[1:4]
try:
    d=json.loads(open("gpurun_out/sw3.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(w, "seg_first", a, "ratio", b, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()}, "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e: print("parse fail", w, a, b, e, open("gpurun_out/sw3.err").read()[-300:])
PY
 done; done
done
