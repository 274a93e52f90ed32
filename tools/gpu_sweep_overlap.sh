#!/bin/bash
mkdir -p gpurun_out
run() {
  env "$@" timeout 600 python bench.py --workload cfg2 --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/sw4.json 2> gpurun_out/sw4.err
  python - "$*" <<'PY'
import json, sys
try:
    d=json.loads(open("gpurun_out/sw4.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "| value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()})
except Exception as e: print("parse fail", sys.argv[1], e, open("gpurun_out/sw4.err").read()[-300:])
PY
}
run A=1
run VB200_K2_SPLIT_MAX_B=32
run VB200_K2_SPLIT_MAX_B=32 VB200_K2_KBOX=1 VB200_K2_SMEM_KB=112
run VB200_K2_SPLIT_MAX_B=32 VB200_K2_KBOX=1 VB200_K2_SMEM_KB=128
run VB200_K2_SPLIT_MAX_B=32 VB200_K2_KBOX=2 VB200_K2_SMEM_KB=128
run VB200_K2_KBOX=1 VB200_K2_SMEM_KB=162
run VB200_K2_KBOX=1 VB200_K2_SMEM_KB=178
