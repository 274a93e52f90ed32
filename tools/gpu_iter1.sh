#!/bin/bash
# iteration check: GPU tests, cfg2 bench, cfg3-b1-s50 launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/it_pytest.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/it_pytest.log
timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/it_cfg2.json 2> gpurun_out/it_cfg2.err; echo "cfg2 rc=$?"; cut -c1-300 gpurun_out/it_cfg2.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/it_cfg2.json").read().strip().splitlines()[-1])
    print("cfg2 value", d["value"], "e2e", d["e2e"]["value"], d["roofline"]["phase_ms_per_step"], d["roofline"]["big_launch"])
except Exception as e: print("parse fail", e)
PY
CMD="python bench.py --workload cfg3-b1-s50 --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/it_cfg3_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/it_launches_cfg3b1s50.csv $CMD > gpurun_out/it_ncu1.log 2>&1
echo "launch list rc=$?"
tail -c 600 gpurun_out/it_cfg3_plain.log
