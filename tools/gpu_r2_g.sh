#!/bin/bash
# Round 2, pass 7: full GPU suite (no -x), cfg5 overflow diagnosis, K1F retune A/B, dense segment schedule sweep.
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f q/s  ms/step %.3f  e2e %.0f  phases/batch %s  timeline %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_batch"].items()}, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (d.get("timeline") or {}).items() if k != "note"}))
except Exception as e:
    print("  no line:", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -25 gpurun_out/r02g_pytest_gpu.log
timeout 600 python tools/debug_overflow.py cfg5-shard > gpurun_out/r02g_debug_cfg5.log 2>&1; echo "debug cfg5 rc=$?"; tail -12 gpurun_out/r02g_debug_cfg5.log
for W in cfg1 cfg3-b1-s1 cfg3-b1-s50; do
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02g_$W.json 2> gpurun_out/r02g_$W.err
  echo "$W rc=$?"; line gpurun_out/r02g_$W.json; tail -2 gpurun_out/r02g_$W.err
done
for SF in "2048 8" "16384 8" "16384 4" "2048 4"; do
  set -- $SF
  VB200_SEG_FIRST=$1 VB200_SEG_RATIO=$2 timeout 600 python bench.py --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02g_cfg4_sf$1_r$2.json 2> gpurun_out/r02g_cfg4_sf$1_r$2.err
  echo "cfg4 seg_first=$1 ratio=$2 rc=$?"; line gpurun_out/r02g_cfg4_sf$1_r$2.json; tail -2 gpurun_out/r02g_cfg4_sf$1_r$2.err
done
for CV in 100; do
  VB200_SPARSE_CARVEOUT=$CV timeout 600 python bench.py --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02g_cfg4_cv$CV.json 2> gpurun_out/r02g_cfg4_cv$CV.err
  echo "cfg4 carveout=$CV rc=$?"; line gpurun_out/r02g_cfg4_cv$CV.json; tail -2 gpurun_out/r02g_cfg4_cv$CV.err
done
for W in cfg2 cfg3-b256-s50 cfg5-shard; do
  timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02g_$W.json 2> gpurun_out/r02g_$W.err
  echo "$W rc=$?"; line gpurun_out/r02g_$W.json; tail -2 gpurun_out/r02g_$W.err
done
