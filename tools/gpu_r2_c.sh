#!/bin/bash
# Round 2, third pass: K3M with bucket tables. GPU suite, workloads, K3M term limit on cfg5, launch lists, ncu.
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
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/r02c_pytest_gpu.log
for W in cfg2 cfg3-b256-s50 cfg4-shard cfg3-b1-s50 cfg3-b1-s1 cfg1; do
  ST=100; [ "$W" = "cfg3-b256-s50" ] && ST=10; [ "$W" = "cfg4-shard" ] && ST=6
  timeout 900 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/r02c_$W.json 2> gpurun_out/r02c_$W.err
  echo "$W rc=$?"; line gpurun_out/r02c_$W.json; tail -2 gpurun_out/r02c_$W.err
done
for MT in 16 8 32 64; do
  VB200_MS_MAX_TERMS=$MT timeout 900 python bench.py --workload cfg5-shard --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_cfg5_mt$MT.json 2> gpurun_out/r02c_cfg5_mt$MT.err
  echo "cfg5-shard max_terms=$MT rc=$?"; line gpurun_out/r02c_cfg5_mt$MT.json; tail -2 gpurun_out/r02c_cfg5_mt$MT.err
done
for W in cfg2 cfg3-b256-s50; do
  CMD="python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/r02c_launches_$W.csv $CMD > gpurun_out/ncu_l_$W.log 2>&1
  echo "launch list $W rc=$?"
done
CMD="python bench.py --workload cfg3-b256-s50 --steps 3 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:vb_ms_score_kernel -s 14 -c 1 -f -o gpurun_out/r02c_prof_ms_cfg3 $CMD > gpurun_out/ncu_ms_cfg3.log 2>&1
echo "ms full cfg3 rc=$?"
