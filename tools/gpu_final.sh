#!/bin/bash
# Round evidence: full GPU suite, smoke, default bench line, reference arm, other workloads,
# ncu launch list + full captures of the three top kernels.
R=${1:-r01}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 900 python bench.py > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${R}_bench_cfg2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/${R}_bench_reference.json
for W in cfg1 cfg3-b1-s1 cfg3-b1-s50 cfg3-b256-s50 cfg4-shard; do
  ST=200; [ "$W" = "cfg3-b256-s50" ] && ST=20; [ "$W" = "cfg4-shard" ] && ST=6
  timeout 900 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu-baseline > gpurun_out/${R}_bench_$W.json 2> gpurun_out/${R}_bench_$W.err; echo "$W rc=$?"
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/${R}_launches_cfg2.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_dense_gemm_kernel -s 8 -c 1 -f -o gpurun_out/${R}_prof_gemm $CMD > gpurun_out/ncu2.log 2>&1
echo "gemm full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_sparse_kernel -s 8 -c 1 -f -o gpurun_out/${R}_prof_sparse $CMD > gpurun_out/ncu3.log 2>&1
echo "sparse full rc=$?"
CMD2="python bench.py --workload cfg3-b256-s50 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_dense_gemm_tiled -s 11 -c 1 -f -o gpurun_out/${R}_prof_tiled $CMD2 > gpurun_out/ncu4.log 2>&1
echo "tiled full rc=$?"
