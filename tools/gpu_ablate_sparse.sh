#!/bin/bash
# timing ablation of vb_sparse_kernel on cfg2 (results are wrong with debug != 0; only the phase times matter)
mkdir -p gpurun_out
for D in 0 4 32; do
  VB200_SPARSE_DEBUG=$D timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - $D <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("debug", sys.argv[1], "ms/step %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()}, "sparse big %.4f" % r["big_launch"]["sparse_ms"])
except Exception as e: print("parse fail", sys.argv[1], e, open("gpurun_out/ab.err").read()[-300:])
PY
done
