#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; env "$@" timeout 300 python bench.py --workload $wl --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read())
print('$wl $*', 'value=%.0f ms=%.3f e2e=%.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items()})"; }
cp voitta-rag_b200/libvoitta_b200.so /tmp/lib_keep.so
for R in 1024 4096; do
cp voitta-rag_b200/libvb_R$R.so voitta-rag_b200/libvoitta_b200.so
run cfg2 VB200_OVERLAP=0 VB200_SEG_FIRST=$R R=$R
run cfg2 VB200_OVERLAP=1 VB200_SEG_FIRST=$R R=$R
done
cp /tmp/lib_keep.so voitta-rag_b200/libvoitta_b200.so
