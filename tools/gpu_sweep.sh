#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_sweep.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_sweep.log
run() { wl=$1; shift; env "$@" timeout 300 python bench.py --workload $wl --steps 300 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('$wl $*', 'value=%.0f ms=%.3f e2e=%.0f (%.3f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()})"; }
run cfg2 X=0
run cfg1 X=0
