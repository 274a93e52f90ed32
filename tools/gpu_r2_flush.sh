#!/bin/bash
# Round 2, pass 32: clean L2 flush (write pass + read pass) — cfg1 / cfg2 / cfg2-g8shard lines
mkdir -p gpurun_out
for W in cfg1 cfg2; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02f_bench_$W.json 2> gpurun_out/r02f_bench_$W.err; echo "$W rc=$?"
  python - <<P
import json
d=json.loads(open("gpurun_out/r02f_bench_$W.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"]/d["config"]["batches_per_step"], round(d["e2e"]["value"]), d["roofline"]["phase_ms_per_batch"], d["roofline"]["frac"])
P
done
