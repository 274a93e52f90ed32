#!/bin/bash
# launch list (ncu gpu__time_duration) of one workload: $1
W=$1
mkdir -p gpurun_out
timeout 600 python bench.py --workload $W --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/lw_$W.json 2> gpurun_out/lw_$W.err; echo "rc=$?"
python - $W <<'PY'
import json, sys
w = sys.argv[1]
d=json.loads(open(f"gpurun_out/lw_{w}.json").read().strip().splitlines()[-1]); r=d["roofline"]
print(w, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()}, "launches/step", d["gpu_launches"]/d["steps"])
PY
CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/lw_launches_$W.csv $CMD > gpurun_out/ncu_lw.log 2>&1
echo "launch list rc=$?"
