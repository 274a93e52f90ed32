#!/bin/bash
# Refresh at HEAD: GPU suite, smoke, the default bench line (reads the regenerated profiles/traffic.json) and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err; echo "default bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_cfg4.json
