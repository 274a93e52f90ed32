#!/bin/bash
# Round 2, pass 5: staged K3M + per-column GEMM epilogue. GPU suite, A/B (staged vs segments), seg-ratio sweep, launch lists.
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
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r02e_pytest_gpu.log
for W in cfg2 cfg3-b256-s50 cfg4; do
  for ST in 1 0; do
    VB200_MS_STAGED=$ST timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02e_${W}_staged$ST.json 2> gpurun_out/r02e_${W}_staged$ST.err
    echo "$W staged=$ST rc=$?"; line gpurun_out/r02e_${W}_staged$ST.json; tail -2 gpurun_out/r02e_${W}_staged$ST.err
  done
done
for R in 8 16; do
  VB200_SEG_RATIO=$R timeout 900 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02e_cfg4_ratio$R.json 2> gpurun_out/r02e_cfg4_ratio$R.err
  echo "cfg4 seg_ratio=$R rc=$?"; line gpurun_out/r02e_cfg4_ratio$R.json; tail -2 gpurun_out/r02e_cfg4_ratio$R.err
done
for W in cfg5-shard cfg3-b1-s50 cfg3-b1-s1 cfg1; do
  timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02e_$W.json 2> gpurun_out/r02e_$W.err
  echo "$W rc=$?"; line gpurun_out/r02e_$W.json; tail -2 gpurun_out/r02e_$W.err
done
for W in cfg2 cfg4 cfg3-b1-s1; do
  CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-api"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/r02e_launches_$W.csv $CMD > gpurun_out/ncu_l_$W.log 2>&1
  echo "launch list $W rc=$?"
done
