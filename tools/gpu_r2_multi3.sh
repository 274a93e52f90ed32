#!/bin/bash
# Round 2: N-GPU default line with the host-time breakdown of the pipelined e2e region
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m_scale_cfg4_g$N.json 2> gpurun_out/r02m_scale_cfg4_g$N.err; echo "cfg4 x$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02m_scale_cfg4_g$N.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], json.dumps(d["e2e"])[:600], d["clocks"])
PY
tail -3 gpurun_out/r02m_scale_cfg4_g$N.err
