#!/bin/bash
# Round 2: BASELINE configs[3] and configs[4] on N GPUs at HEAD (row selection on), launched as the driver launches bench.py
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02n_scale_cfg4_g$N.json 2> gpurun_out/r02n_scale_cfg4_g$N.err; echo "cfg4 x$N rc=$?"
if [ "$N" = "8" ]; then
timeout 300 $TR bench.py --gpus $N --workload cfg5 --steps 4 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r02n_scale_cfg5_g$N.json 2> gpurun_out/r02n_scale_cfg5_g$N.err; echo "cfg5 x$N rc=$?"
fi
python - <<PY
import json
for w in ("cfg4","cfg5"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02n_scale_{w}_g$N.json") if l.startswith("{")][-1])
        print(w, d["value"], d["ms_per_step"], json.dumps(d["e2e"])[:300], d["clocks"], d["roofline"]["launch"][:200])
    except Exception as e: print(w, "parse", e)
PY
tail -3 gpurun_out/r02n_scale_cfg4_g$N.err
