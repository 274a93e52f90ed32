#!/bin/bash
# N-GPU runs (N = $1): cfg2 weak-scaling line and the cfg4 shape (12.5M rows per GPU, batch 1024, top-100)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/mg_cfg2_g$N.json 2> gpurun_out/mg_cfg2_g$N.err; echo "cfg2 x$N rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/mg_cfg4_g$N.json 2> gpurun_out/mg_cfg4_g$N.err; echo "cfg4 x$N rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for w in ("cfg2", "cfg4"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/mg_{w}_g{n}.json").read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
        print(w, "x"+n, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()}, r["bound"], round(r["frac"],3))
    except Exception as e:
        print("parse fail", w, e); print(open(f"gpurun_out/mg_{w}_g{n}.err").read()[-1500:])
PY
