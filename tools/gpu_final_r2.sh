#!/bin/bash
# Round 2 evidence: GPU suite, smoke, default bench line + reference arm, one line per BASELINE config, bounded ncu
# launch lists, --set full captures of the dominant kernels, compute-sanitizer.  Everything lands in gpurun_out/r02_*;
# tools/summarize_profiles_r2.py (run in the dev container afterwards) turns it into the tracked files under profiles/.
R=r02
mkdir -p gpurun_out
SEARCH='regex:vb_(ms_|mh_|dense|compact|fuse|mask|sparse_kernel|sparse_plan|sparse_delta|slice|prep|init_lists|rowsel)'
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg4.json 2> gpurun_out/${R}_bench_cfg4.err; echo "default bench rc=$?"; cut -c1-300 gpurun_out/${R}_bench_cfg4.json
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/${R}_bench_reference.json
timeout 900 python bench.py --workload cfg2 --steps 20 --warmup 5 > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; echo "cfg2 rc=$?"
for W in cfg1 cfg3-b1-s1 cfg3-b1-s50 cfg3-b256-s1 cfg3-b256-s50 cfg5-shard; do
  timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_$W.json 2> gpurun_out/${R}_bench_$W.err; echo "$W rc=$?"
done
# per-launch device times (bounded: a step is a stream of batches)
for W in cfg4 cfg2 cfg3-b1-s1 cfg5-shard; do
  C=160; [ "$W" = "cfg5-shard" ] && C=120
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$SEARCH" -c $C --csv --log-file gpurun_out/${R}_launches_$W.csv \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/ncu_l_$W.log 2>&1
  echo "launch list $W rc=$?"
done
# full captures: name  workload  kernel-regex  launches-to-skip.  Only the raw-metric CSV travels back (gpurun_out is
# capped at 64 MiB and a --set full report is 5-15 MB); the two dominant kernels keep their .ncu-rep as well.
cap() {
  timeout 900 ncu --set full --clock-control none -k regex:$3 -s $4 -c 1 -f -o gpurun_out/${R}_prof_$1 \
    python bench.py --workload $2 --steps 1 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 ($2 $3 skip $4) rc=$?"
  ncu -i gpurun_out/${R}_prof_$1.ncu-rep --page raw --csv > gpurun_out/${R}_prof_$1.csv 2>/dev/null
  [ "$5" = "keep" ] || rm -f gpurun_out/${R}_prof_$1.ncu-rep
  tail -c 2000 gpurun_out/ncu_$1.log > gpurun_out/ncu_$1.tail; rm -f gpurun_out/ncu_$1.log
}
cap k2t cfg4 vb_dense_gemm_tiled_kernel 10 keep
cap k3m cfg4 vb_ms_score_kernel 9 keep
cap sel cfg4 vb_rowsel_gather_kernel 4
cap mask cfg4 vb_mask_kernel 3
cap compact cfg4 vb_compact_kernel 40
cap k1 cfg3-b1-s50 vb_dense_scan_kernel 8
cap k1f cfg1 vb_dense_scan1_kernel 5
cap k3 cfg5-shard vb_sparse_kernel 6
cap k2 cfg2 vb_dense_gemm_kernel 8
rm -f gpurun_out/ncu_l_*.log
# bounds-checked build (compute-sanitizer is closed on the pool): all-kernel workload + GPU suite with every VB_CHECK armed
bash tools/gpu_sanitize.sh ${R}
du -sh gpurun_out
