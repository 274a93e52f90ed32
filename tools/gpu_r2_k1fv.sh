#!/bin/bash
# Round 2, pass 36: K1F with launch bounds that force 5 / 6 / 8 resident CTAs per SM (fewer registers, some spills) against
# the default build (62 registers at d = 384: 4 CTAs per SM) — cfg1 lines
mkdir -p gpurun_out
for L in default k1f5 k1f6 k1f8 default; do
  if [ $L = default ]; then unset VB200_LIB; else export VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_$L.so; fi
  timeout 120 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02k1fv_$L.json 2> gpurun_out/r02k1fv_$L.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02k1fv_$L.json").read().strip().splitlines()[-1])
    print("$L", round(d["value"]), round(d["ms_per_step"]/d["config"]["batches_per_step"],5), d["roofline"]["phase_ms_per_batch"]["dense"])
except Exception as e: print("$L", "parse", e)
P
done
