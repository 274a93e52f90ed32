#!/bin/bash
# Round evidence: full GPU suite, smoke, default bench line, reference arm, ncu launch list + full captures.
R=${1:-r01}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${R}_smoke.log
timeout 900 python bench.py > gpurun_out/${R}_bench_cfg2.json 2> gpurun_out/${R}_bench_cfg2.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${R}_bench_cfg2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/${R}_bench_reference.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_ --csv --log-file gpurun_out/${R}_launches_cfg2.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_dense_gemm -s 8 -c 1 -o gpurun_out/${R}_prof_gemm $CMD > gpurun_out/ncu2.log 2>&1
echo "gemm full rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vb_sparse_kernel -s 8 -c 1 -o gpurun_out/${R}_prof_sparse $CMD > gpurun_out/ncu3.log 2>&1
echo "sparse full rc=$?"
