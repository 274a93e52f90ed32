#!/bin/bash
# Round 2, pass 9: bounds-checked build over the all-kernel workload + GPU suite; full suite on the release build;
# quick lines after the last changes (compaction fast path, fused list bookkeeping).
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f q/s  ms/batch %.4f  e2e %.0f  phases/batch %s  roofline %s %s frac %.3f traffic %s" % (d["value"], d["ms_per_step"] / d["config"]["batches_per_step"], d["e2e"]["value"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_batch"].items()}, d["roofline"]["kernel"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"]["traffic"]))
except Exception as e:
    print("  no line:", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02i_pytest_gpu.log
bash tools/gpu_sanitize.sh r02
for W in cfg1 cfg3-b1-s1 cfg2 cfg4; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02i_$W.json 2> gpurun_out/r02i_$W.err
  echo "$W rc=$?"; line gpurun_out/r02i_$W.json; tail -2 gpurun_out/r02i_$W.err
done
