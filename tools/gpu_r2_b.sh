#!/bin/bash
# Round 2, second pass: K3M work-unit size sweep, per-kernel launch lists, ncu capture of vb_ms_score_kernel.
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f q/s  ms/step %.3f  phases %s" % (d["value"], d["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["phase_ms_per_step"].items()}))
except Exception as e:
    print("  no line:", e)
PY
}
for W in cfg2 cfg3-b256-s50; do
  ST=100; [ "$W" = "cfg3-b256-s50" ] && ST=10
  for CH in 256 512 1024 4096; do
    VB200_MS_CHUNK=$CH timeout 600 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/r02b_${W}_ch$CH.json 2> gpurun_out/r02b_${W}_ch$CH.err
    echo "$W chunk=$CH rc=$?"; line gpurun_out/r02b_${W}_ch$CH.json
  done
done
for W in cfg4-shard cfg5-shard; do
  ST=6; [ "$W" = "cfg5-shard" ] && ST=3
  timeout 900 python bench.py --workload $W --steps $ST --warmup 3 --no-cpu-baseline > gpurun_out/r02b_$W.json 2> gpurun_out/r02b_$W.err
  echo "$W rc=$?"; line gpurun_out/r02b_$W.json; tail -2 gpurun_out/r02b_$W.err
done
for W in cfg2 cfg3-b256-s50; do
  CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/r02b_launches_$W.csv $CMD > gpurun_out/ncu_l_$W.log 2>&1
  echo "launch list $W rc=$?"
done
CMD="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:vb_ms_score_kernel -s 9 -c 1 -f -o gpurun_out/r02b_prof_ms_cfg2 $CMD > gpurun_out/ncu_ms_cfg2.log 2>&1
echo "ms full cfg2 rc=$?"
CMD="python bench.py --workload cfg3-b256-s50 --steps 3 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:vb_ms_score_kernel -s 14 -c 1 -f -o gpurun_out/r02b_prof_ms_cfg3 $CMD > gpurun_out/ncu_ms_cfg3.log 2>&1
echo "ms full cfg3 rc=$?"
