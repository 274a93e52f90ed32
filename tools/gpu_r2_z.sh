#!/bin/bash
# Round 2, pass 25: does the 50 ms nvidia-smi sampler perturb the timed region?  (value 22.2 ms vs 19.9 ms timeline span at cfg4)
mkdir -p gpurun_out
for P in 50 250 1000 50; do
  timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu-baseline --no-api --clock-period-ms $P > gpurun_out/r02z_bench_cfg4_p$P.json 2> gpurun_out/r02z_bench_cfg4_p$P.err; echo "period $P rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02z_bench_cfg4_p$P.json").read().strip().splitlines()[-1])
print($P, round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"], d["timeline"]["span_ms"])
PY
done
