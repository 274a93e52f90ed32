#!/bin/bash
# Full GPU suite + smoke + first bench lines.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --workload tiny --steps 50 --warmup 5 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "bench tiny rc=$?"; tail -c 1500 gpurun_out/bench_tiny.json; tail -5 gpurun_out/bench_tiny.err
timeout 900 python bench.py --steps 300 --warmup 10 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"; cat gpurun_out/bench_cfg2.json; tail -5 gpurun_out/bench_cfg2.err
timeout 600 python bench.py --steps 300 --warmup 10 --dense-path 1 --no-cpu-baseline > gpurun_out/bench_cfg2_k1.json 2> gpurun_out/bench_cfg2_k1.err; echo "bench cfg2 k1 rc=$?"; cat gpurun_out/bench_cfg2_k1.json; tail -3 gpurun_out/bench_cfg2_k1.err
