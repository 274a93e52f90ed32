#!/bin/bash
# Round 2, pass 4: GPU suite (incremental index), default bench line (cfg4 weak form) + reference arm, cfg2 line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r02d_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02d_smoke.log
S=$(date +%s); timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - S )) s"; cut -c1-1500 gpurun_out/r02d_bench_default.json; tail -3 gpurun_out/r02d_bench_default.err
S=$(date +%s); timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err; echo "reference rc=$? wall $(( $(date +%s) - S )) s"; cut -c1-600 gpurun_out/r02d_bench_reference.json; tail -3 gpurun_out/r02d_bench_reference.err
timeout 900 python bench.py --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_cfg2.json 2> gpurun_out/r02d_bench_cfg2.err; echo "cfg2 rc=$?"; cut -c1-2500 gpurun_out/r02d_bench_cfg2.json; tail -3 gpurun_out/r02d_bench_cfg2.err
