#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; timeout 900 python bench.py --workload $wl --steps $1 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$wl.err | tee gpurun_out/bench_$wl.json | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('$wl', 'value=%.1f ms=%.3f e2e=%.1f path=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['dense_path']), {k: round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()}, 'frac=%.3f' % d['roofline']['frac'], d['roofline']['kernel'])"; tail -2 gpurun_out/bench_$wl.err; }
run cfg1 200
run cfg3-b1-s50 30
run cfg3-b1-s1 30
run cfg3-b256-s50 10
nvidia-smi --query-gpu=memory.used --format=csv
