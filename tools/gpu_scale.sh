#!/bin/bash
# cfg2 bench line at N GPUs exactly as the driver launches it (N = $1), plus cfg4 when N = 8
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 500 --warmup 10 > gpurun_out/r01_scale_cfg2_g$N.json 2> gpurun_out/r01_scale_cfg2_g$N.err; echo "cfg2 x$N rc=$?"
if [ "$N" = "8" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r01_scale_cfg4_g$N.json 2> gpurun_out/r01_scale_cfg4_g$N.err; echo "cfg4 x$N rc=$?"
fi
python - $N <<'PY'
import json, sys, os
n = sys.argv[1]
for w in ("cfg2", "cfg4"):
    f = f"gpurun_out/r01_scale_{w}_g{n}.json"
    if not os.path.exists(f): continue
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
        print(w, "x"+n, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()})
    except Exception as e:
        print("parse fail", w, e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
