#!/bin/bash
# Round 2, pass 11: K1F v2 (one wave, in-kernel merge, multi-group units) — GPU suite, then B = 1 lines with k1f = 1 / 2
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f q/s  ms/batch %.4f  e2e %.0f  launches/batch %.1f phases/batch %s" % (d["value"], d["ms_per_step"] / d["config"]["batches_per_step"], d["e2e"]["value"], d["gpu_launches"] / d["steps"] / d["config"]["batches_per_step"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_batch"].items()}))
except Exception as e:
    print("  no line:", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/r02k_pytest_gpu.log
for W in cfg1 cfg3-b1-s1 cfg3-b1-s50; do
  for K in 1 2; do
    timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-api --opt k1f=$K > gpurun_out/r02k_${W}_k$K.json 2> gpurun_out/r02k_${W}_k$K.err
    echo "$W k1f=$K rc=$?"; line gpurun_out/r02k_${W}_k$K.json; tail -2 gpurun_out/r02k_${W}_k$K.err
  done
done
