#!/bin/bash
# A/B of the threshold exchange at N GPUs on cfg2 (weak scaling shape)
N=${1:-8}
mkdir -p gpurun_out
for S in 1 0; do
VB200_SHARE_TAU=$S python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$S bench.py --gpus $N --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/share_$S.json 2> gpurun_out/share_$S.err; echo "share=$S rc=$?"
python - $S <<'PY'
import json, sys
s = sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/share_{s}.json").read().strip().splitlines() if l.startswith("{")][-1]); r=d["roofline"]
    print("share", s, "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in r["phase_ms_per_step"].items()})
except Exception as e:
    print("parse fail", e); print(open(f"gpurun_out/share_{s}.err").read()[-1500:])
PY
done
