#!/bin/bash
# Round 2, pass 6: K1F + adaptive stages; two-stream overlap experiments (timeline, carveout); sanitizer.
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
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r02f_pytest_gpu.log
for W in cfg1 cfg3-b1-s1 cfg3-b1-s50; do
  for K in 1 0; do
    VB200_K1F=$K timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02f_${W}_k1f$K.json 2> gpurun_out/r02f_${W}_k1f$K.err
    echo "$W k1f=$K rc=$?"; line gpurun_out/r02f_${W}_k1f$K.json; tail -2 gpurun_out/r02f_${W}_k1f$K.err
  done
done
for CV in -1 100 50; do
  VB200_SPARSE_CARVEOUT=$CV timeout 900 python bench.py --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02f_cfg4_cv$CV.json 2> gpurun_out/r02f_cfg4_cv$CV.err
  echo "cfg4 carveout=$CV rc=$?"; line gpurun_out/r02f_cfg4_cv$CV.json; tail -2 gpurun_out/r02f_cfg4_cv$CV.err
done
VB200_K2T_STAGES=3 VB200_SPARSE_CARVEOUT=100 timeout 900 python bench.py --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02f_cfg4_st3.json 2> gpurun_out/r02f_cfg4_st3.err
echo "cfg4 K2T 3 stages + carveout 100 rc=$?"; line gpurun_out/r02f_cfg4_st3.json
VB200_OVERLAP=0 timeout 900 python bench.py --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02f_cfg4_noov.json 2> gpurun_out/r02f_cfg4_noov.err
echo "cfg4 overlap=0 rc=$?"; line gpurun_out/r02f_cfg4_noov.json
bash tools/gpu_sanitize.sh r02f
