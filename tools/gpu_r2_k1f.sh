#!/bin/bash
# Round 2, pass 31: K1F merge with the final-threshold prefilter: GPU suite + cfg1 / cfg3-b1-s1 lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02k_pytest_gpu.log
for W in cfg1 cfg2; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_bench_$W.json 2> gpurun_out/r02k_bench_$W.err; echo "$W rc=$?"
  python - <<P
import json
d=json.loads(open("gpurun_out/r02k_bench_$W.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"]/d["config"]["batches_per_step"], round(d["e2e"]["value"]), d["roofline"]["phase_ms_per_batch"], d["e2e_api"])
P
done
